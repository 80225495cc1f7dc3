//! Two dimensional interpolation along the first two axes: `Interp2D`, `Interp2DBuilder`, the
//! strategy traits and `Bilinear` (K4 `interp2d_bilinear_kernel` behind `ndi_interp2d_bilinear`).
use std::{ffi::c_void, fmt::Debug};

use ndarray::{Array, Array1, ArrayBase, ArrayView, ArrayViewMut, ArrayViewMut1, Axis, Data, DimAdd, Dimension, Ix1, Ix2, OwnedRepr, RemoveAxis};
use num_traits::cast;

use crate::{
    ffi,
    vector_extensions::{Monotonic, VectorExtensions},
    BuilderError, InterpolateError, NdiElem,
};

#[derive(Debug)]
pub struct DeviceTable2D(pub(crate) *mut ffi::ndi_interp2d);
unsafe impl Send for DeviceTable2D {}
unsafe impl Sync for DeviceTable2D {}
impl Drop for DeviceTable2D {
    fn drop(&mut self) {
        unsafe { ffi::ndi_interp2d_destroy(self.0) };
    }
}

type Smaller2<D> = <<D as Dimension>::Smaller as Dimension>::Smaller;

pub trait Interp2DStrategyBuilder<Sd, Sx, Sy, D>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    D: Dimension + RemoveAxis,
    D::Smaller: RemoveAxis,
{
    const MINIMUM_DATA_LENGHT: usize;
    type FinishedStrat: Interp2DStrategy<Sd, Sx, Sy, D>;
    fn build(self, x: &ArrayBase<Sx, Ix1>, y: &ArrayBase<Sy, Ix1>, data: &ArrayBase<Sd, D>) -> Result<Self::FinishedStrat, BuilderError>;
}

pub trait Interp2DStrategy<Sd, Sx, Sy, D>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    D: Dimension + RemoveAxis,
    D::Smaller: RemoveAxis,
    Self: Sized,
{
    /// Interpolate at position `(x, y)` into `target` (shape = data shape without axes 0 and 1).
    fn interp_into(&self, interpolator: &Interp2D<Sd, Sx, Sy, D, Self>, target: ArrayViewMut<'_, Sd::Elem, <D::Smaller as Dimension>::Smaller>, x: Sx::Elem, y: Sy::Elem) -> Result<(), InterpolateError>;

    /// the batch loop; default = the reference's loop over `interp_into`, `Bilinear` = one launch
    fn interp_batch_into(&self, interpolator: &Interp2D<Sd, Sx, Sy, D, Self>, xs: &[Sd::Elem], ys: &[Sd::Elem], out: &mut [Sd::Elem]) -> Result<(), InterpolateError> {
        let row_dim = interpolator.data.raw_dim().remove_axis(Axis(0)).remove_axis(Axis(0));
        let w = row_dim.size();
        for (i, (&x, &y)) in xs.iter().zip(ys).enumerate() {
            let row = ArrayViewMut::from_shape(row_dim.clone(), &mut out[i * w..(i + 1) * w]).unwrap_or_else(|_| unreachable!());
            self.interp_into(interpolator, row, x, y)?;
        }
        Ok(())
    }
}

/// Bilinear strategy
#[derive(Debug)]
pub struct Bilinear {
    extrapolate: bool,
}

impl Bilinear {
    pub fn new() -> Self {
        Bilinear { extrapolate: false }
    }
    pub fn extrapolate(mut self, yes: bool) -> Self {
        self.extrapolate = yes;
        self
    }
}

impl Default for Bilinear {
    fn default() -> Self {
        Self::new()
    }
}

impl<Sd, Sx, Sy, D> Interp2DStrategyBuilder<Sd, Sx, Sy, D> for Bilinear
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    D: Dimension + RemoveAxis,
    D::Smaller: RemoveAxis,
{
    const MINIMUM_DATA_LENGHT: usize = 2;
    type FinishedStrat = Self;
    fn build(self, _x: &ArrayBase<Sx, Ix1>, _y: &ArrayBase<Sy, Ix1>, _data: &ArrayBase<Sd, D>) -> Result<Self, BuilderError> {
        Ok(self)
    }
}

impl<Sd, Sx, Sy, D> Interp2DStrategy<Sd, Sx, Sy, D> for Bilinear
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    D: Dimension + RemoveAxis,
    D::Smaller: RemoveAxis,
{
    fn interp_into(&self, interpolator: &Interp2D<Sd, Sx, Sy, D, Self>, mut target: ArrayViewMut<'_, Sd::Elem, <D::Smaller as Dimension>::Smaller>, x: Sx::Elem, y: Sy::Elem) -> Result<(), InterpolateError> {
        match target.as_slice_mut() {
            Some(out) => self.interp_batch_into(interpolator, &[x], &[y], out),
            None => {
                let mut scratch = Array::<Sd::Elem, _>::zeros(target.raw_dim());
                let res = self.interp_batch_into(interpolator, &[x], &[y], scratch.as_slice_mut().unwrap_or_else(|| unreachable!()));
                target.assign(&scratch);
                res
            }
        }
    }

    fn interp_batch_into(&self, interpolator: &Interp2D<Sd, Sx, Sy, D, Self>, xs: &[Sd::Elem], ys: &[Sd::Elem], out: &mut [Sd::Elem]) -> Result<(), InterpolateError> {
        let (mut first_bad, mut axis) = (-1i64, -1i32);
        let st = unsafe {
            ffi::ndi_interp2d_bilinear(
                interpolator.table.0,
                xs.as_ptr() as *const c_void,
                ys.as_ptr() as *const c_void,
                xs.len() as i64,
                self.extrapolate as i32,
                out.as_mut_ptr() as *mut c_void,
                &mut first_bad,
                &mut axis,
            )
        };
        match st {
            ffi::NDI_OK => Ok(()),
            ffi::NDI_OUT_OF_BOUNDS if axis == 0 => Err(InterpolateError::OutOfBounds(format!("x = {:?} is not in range", xs[first_bad as usize]))),
            ffi::NDI_OUT_OF_BOUNDS => Err(InterpolateError::OutOfBounds(format!("y = {:?} is not in range", ys[first_bad as usize]))),
            ffi::NDI_NAN_QUERY => unimplemented!("failed to convert NaN to usize"),
            _ => panic!("device evaluation failed ({st}): {}", ffi::last_error()),
        }
    }
}

/// Two dimensional interpolator
#[derive(Debug)]
pub struct Interp2D<Sd, Sx, Sy, D, Strat>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    D: Dimension,
{
    pub(crate) x: ArrayBase<Sx, Ix1>,
    pub(crate) y: ArrayBase<Sy, Ix1>,
    pub(crate) data: ArrayBase<Sd, D>,
    pub(crate) strategy: Strat,
    pub(crate) table: DeviceTable2D,
}

/// Create and configure a [Interp2D] interpolator.
#[derive(Debug)]
pub struct Interp2DBuilder<Sd, Sx, Sy, D, Strat>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    D: Dimension,
{
    x: ArrayBase<Sx, Ix1>,
    y: ArrayBase<Sy, Ix1>,
    data: ArrayBase<Sd, D>,
    strategy: Strat,
}

impl<Sd, D> Interp2D<Sd, OwnedRepr<Sd::Elem>, OwnedRepr<Sd::Elem>, D, Bilinear>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    D: Dimension,
{
    /// Get the [Interp2DBuilder]
    pub fn builder(data: ArrayBase<Sd, D>) -> Interp2DBuilder<Sd, OwnedRepr<Sd::Elem>, OwnedRepr<Sd::Elem>, D, Bilinear> {
        Interp2DBuilder::new(data)
    }
}

impl<Sd, Sx, Sy, Strat> Interp2D<Sd, Sx, Sy, Ix2, Strat>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    Strat: Interp2DStrategy<Sd, Sx, Sy, Ix2>,
{
    /// interpolation at one point when the data dimension is `Ix2`
    pub fn interp_scalar(&self, x: Sx::Elem, y: Sy::Elem) -> Result<Sd::Elem, InterpolateError> {
        let mut buffer = [cast::<f64, Sd::Elem>(0.0).unwrap_or_else(|| unimplemented!())];
        let view = ArrayViewMut1::from(buffer.as_mut_slice()).remove_axis(Axis(0));
        self.strategy.interp_into(self, view, x, y).map(|_| buffer[0])
    }
}

impl<Sd, Sx, Sy, D, Strat> Interp2D<Sd, Sx, Sy, D, Strat>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    D: Dimension + RemoveAxis,
    D::Smaller: RemoveAxis,
    Strat: Interp2DStrategy<Sd, Sx, Sy, D>,
{
    pub fn interp(&self, x: Sx::Elem, y: Sy::Elem) -> Result<Array<Sd::Elem, Smaller2<D>>, InterpolateError> {
        let mut target = Array::zeros(self.data.raw_dim().remove_axis(Axis(0)).remove_axis(Axis(0)));
        self.strategy.interp_into(self, target.view_mut(), x, y).map(|_| target)
    }

    pub fn interp_into(&self, x: Sx::Elem, y: Sy::Elem, buffer: ArrayViewMut<'_, Sd::Elem, Smaller2<D>>) -> Result<(), InterpolateError> {
        self.strategy.interp_into(self, buffer, x, y)
    }

    /// # panics
    /// when `xs.shape() != ys.shape()`
    pub fn interp_array<Sqx, Sqy, Dq>(&self, xs: &ArrayBase<Sqx, Dq>, ys: &ArrayBase<Sqy, Dq>) -> Result<Array<Sd::Elem, <Dq as DimAdd<Smaller2<D>>>::Output>, InterpolateError>
    where
        Sqx: Data<Elem = Sd::Elem>,
        Sqy: Data<Elem = Sd::Elem>,
        Dq: Dimension + DimAdd<Smaller2<D>>,
    {
        assert!(xs.shape() == ys.shape(), "`xs.shape()` and `ys.shape()` do not match");
        let mut shape = <Dq as DimAdd<Smaller2<D>>>::Output::zeros(xs.ndim() + self.data.ndim() - 2);
        for (dst, src) in shape.slice_mut().iter_mut().zip(xs.shape().iter().chain(self.data.shape()[2..].iter())) {
            *dst = *src;
        }
        let mut zs = Array::zeros(shape);
        self.interp_array_into(xs, ys, zs.view_mut()).map(|_| zs)
    }

    /// # panics
    /// when `xs.shape() != ys.shape()` or the buffer has the wrong shape
    pub fn interp_array_into<Sqx, Sqy, Dq>(&self, xs: &ArrayBase<Sqx, Dq>, ys: &ArrayBase<Sqy, Dq>, mut buffer: ArrayViewMut<'_, Sd::Elem, <Dq as DimAdd<Smaller2<D>>>::Output>) -> Result<(), InterpolateError>
    where
        Sqx: Data<Elem = Sd::Elem>,
        Sqy: Data<Elem = Sd::Elem>,
        Dq: Dimension + DimAdd<Smaller2<D>>,
    {
        assert!(xs.shape() == ys.shape(), "`xs.shape()` and `ys.shape()` do not match");
        let expect: Vec<usize> = xs.shape().iter().chain(self.data.shape()[2..].iter()).copied().collect();
        assert!(buffer.shape() == expect.as_slice(), "expected: {:?}, got: {:?}", expect, buffer.shape());
        let (qx, qy) = (xs.as_standard_layout(), ys.as_standard_layout());
        let (qx, qy) = (qx.as_slice().unwrap_or_else(|| unreachable!()), qy.as_slice().unwrap_or_else(|| unreachable!()));
        match buffer.as_slice_mut() {
            Some(out) => self.strategy.interp_batch_into(self, qx, qy, out),
            None => {
                let mut scratch = Array::<Sd::Elem, _>::zeros(buffer.raw_dim());
                let res = self.strategy.interp_batch_into(self, qx, qy, scratch.as_slice_mut().unwrap_or_else(|| unreachable!()));
                buffer.assign(&scratch);
                res
            }
        }
    }

    /// Create a interpolator without any data validation.
    pub fn new_unchecked(x: ArrayBase<Sx, Ix1>, y: ArrayBase<Sy, Ix1>, data: ArrayBase<Sd, D>, strategy: Strat) -> Self {
        let table = upload(&x, &y, &data, ffi::NDI_ASSUME_VALID).unwrap_or_else(|e| panic!("{e}"));
        Interp2D { x, y, data, strategy, table }
    }

    /// get `(x, y, data)` coordinate at the given index
    pub fn index_point(&self, x_idx: usize, y_idx: usize) -> (Sx::Elem, Sx::Elem, ArrayView<'_, Sd::Elem, Smaller2<D>>) {
        (self.x[x_idx], self.y[y_idx], self.data.index_axis(Axis(0), x_idx).index_axis_move(Axis(0), y_idx))
    }

    pub fn get_index_left_of(&self, x: Sx::Elem, y: Sy::Elem) -> (usize, usize) {
        (self.x.get_lower_index(x), self.y.get_lower_index(y))
    }

    pub fn is_in_x_range(&self, x: Sx::Elem) -> bool {
        self.x[0] <= x && x <= self.x[self.x.len() - 1]
    }
    pub fn is_in_y_range(&self, y: Sy::Elem) -> bool {
        self.y[0] <= y && y <= self.y[self.y.len() - 1]
    }
}

fn upload<Sd, Sx, Sy, D>(x: &ArrayBase<Sx, Ix1>, y: &ArrayBase<Sy, Ix1>, data: &ArrayBase<Sd, D>, flags: u32) -> Result<DeviceTable2D, BuilderError>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    D: Dimension,
{
    // views go up as they lie in memory; the library makes them dense on the device
    let shape: Vec<i64> = data.shape().iter().map(|&s| s as i64).collect();
    let strides: Vec<i64> = data.strides().iter().map(|&s| s as i64).collect();
    let mut handle = std::ptr::null_mut();
    let st = unsafe {
        ffi::ndi_interp2d_create_strided(
            <Sd::Elem as NdiElem>::DTYPE,
            x.as_ptr() as *const c_void,
            x.len() as i64,
            x.strides()[0] as i64,
            y.as_ptr() as *const c_void,
            y.len() as i64,
            y.strides()[0] as i64,
            data.as_ptr() as *const c_void,
            shape.len() as i32,
            shape.as_ptr(),
            strides.as_ptr(),
            flags,
            &mut handle,
        )
    };
    match st {
        ffi::NDI_OK => Ok(DeviceTable2D(handle)),
        ffi::NDI_NOT_MONOTONIC => Err(BuilderError::Monotonic(ffi::last_error())),
        _ => panic!("ndi_interp2d_create failed ({st}): {}", ffi::last_error()),
    }
}

impl<Sd, D> Interp2DBuilder<Sd, OwnedRepr<Sd::Elem>, OwnedRepr<Sd::Elem>, D, Bilinear>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    D: Dimension,
{
    pub fn new(data: ArrayBase<Sd, D>) -> Self {
        let axis = |len: usize| Array1::from_iter((0..len).map(|i| cast(i).unwrap_or_else(|| unimplemented!("casting from usize to a number should always work"))));
        let (x, y) = (axis(data.shape()[0]), axis(data.shape()[1]));
        Interp2DBuilder { x, y, data, strategy: Bilinear::new() }
    }
}

impl<Sd, Sx, Sy, D, Strat> Interp2DBuilder<Sd, Sx, Sy, D, Strat>
where
    Sd: Data,
    Sd::Elem: NdiElem,
    Sx: Data<Elem = Sd::Elem>,
    Sy: Data<Elem = Sd::Elem>,
    D: Dimension + RemoveAxis,
    D::Smaller: RemoveAxis,
    Strat: Interp2DStrategyBuilder<Sd, Sx, Sy, D>,
{
    pub fn strategy<NewStrat: Interp2DStrategyBuilder<Sd, Sx, Sy, D>>(self, strategy: NewStrat) -> Interp2DBuilder<Sd, Sx, Sy, D, NewStrat> {
        Interp2DBuilder { x: self.x, y: self.y, data: self.data, strategy }
    }

    pub fn x<NewSx: Data<Elem = Sd::Elem>>(self, x: ArrayBase<NewSx, Ix1>) -> Interp2DBuilder<Sd, NewSx, Sy, D, Strat>
    where
        Strat: Interp2DStrategyBuilder<Sd, NewSx, Sy, D>,
    {
        Interp2DBuilder { x, y: self.y, data: self.data, strategy: self.strategy }
    }

    pub fn y<NewSy: Data<Elem = Sd::Elem>>(self, y: ArrayBase<NewSy, Ix1>) -> Interp2DBuilder<Sd, Sx, NewSy, D, Strat>
    where
        Strat: Interp2DStrategyBuilder<Sd, Sx, NewSy, D>,
    {
        Interp2DBuilder { x: self.x, y, data: self.data, strategy: self.strategy }
    }

    /// Validate the input and create the configured [`Interp2D`].
    /// Check order as in the reference: ndim, lengths, length match, THEN monotonicity (x before y).
    pub fn build(self) -> Result<Interp2D<Sd, Sx, Sy, D, Strat::FinishedStrat>, BuilderError> {
        use BuilderError::*;
        let Interp2DBuilder { x, y, data, strategy } = self;
        if data.ndim() < 2 {
            return Err(ShapeError("data dimension needs to be at least 2".into()));
        }
        for axis in 0..2 {
            if data.shape()[axis] < Strat::MINIMUM_DATA_LENGHT {
                return Err(NotEnoughData(format!(
                    "The {axis}-dimension has not enough data for the chosen interpolation strategy. Provided: {}, Reqired: {}",
                    data.shape()[axis],
                    Strat::MINIMUM_DATA_LENGHT
                )));
            }
        }
        if x.len() != data.shape()[0] {
            return Err(ShapeError(format!("Lenghts of x-axis and data-0-axis need to match. Got x: {}, data-0: {}", x.len(), data.shape()[0])));
        }
        if y.len() != data.shape()[1] {
            return Err(ShapeError(format!("Lenghts of y-axis and data-1-axis need to match. Got y: {}, data-1: {}", y.len(), data.shape()[1])));
        }
        if !matches!(x.monotonic_prop(), Monotonic::Rising { strict: true }) {
            return Err(Monotonic("The x-axis needs to be strictly monotonic rising".into()));
        }
        if !matches!(y.monotonic_prop(), Monotonic::Rising { strict: true }) {
            return Err(Monotonic("The y-axis needs to be strictly monotonic rising".into()));
        }
        let table = upload(&x, &y, &data, ffi::NDI_ASSUME_VALID)?;
        let strategy = strategy.build(&x, &y, &data)?;
        Ok(Interp2D { x, y, data, strategy, table })
    }
}
