//! `VectorExtensions` on the device (K1 grid classification, K2 lower-index search).
use std::ffi::c_void;

use ndarray::{ArrayBase, Data, Ix1};

use crate::{ffi, NdiElem};

/// Helper methods for one dimensional numeric arrays
pub trait VectorExtensions<T> {
    /// get the monotonic property of the vector
    fn monotonic_prop(&self) -> Monotonic;
    /// Get the index of the next lower value inside the vector (never the last index).
    fn get_lower_index(&self, x: T) -> usize;
}

/// Describes the monotonic property of a vector
#[derive(Debug)]
pub enum Monotonic {
    Rising { strict: bool },
    Falling { strict: bool },
    NotMonotonic,
}

impl<S> VectorExtensions<S::Elem> for ArrayBase<S, Ix1>
where
    S: Data,
    S::Elem: NdiElem,
{
    fn monotonic_prop(&self) -> Monotonic {
        if self.len() <= 1 {
            return Monotonic::NotMonotonic;
        }
        let mut prop = 0i32;
        // the view may be strided or reversed: pass the first logical element and the stride
        let st = unsafe {
            ffi::ndi_monotonic_prop(
                <S::Elem as NdiElem>::DTYPE,
                self.as_ptr() as *const c_void,
                self.len() as i64,
                self.strides()[0] as i64,
                &mut prop,
            )
        };
        assert!(st == ffi::NDI_OK, "ndi_monotonic_prop: {}", ffi::last_error());
        match prop {
            1 => Monotonic::Rising { strict: true },
            2 => Monotonic::Rising { strict: false },
            3 => Monotonic::Falling { strict: true },
            4 => Monotonic::Falling { strict: false },
            _ => Monotonic::NotMonotonic,
        }
    }

    fn get_lower_index(&self, x: S::Elem) -> usize {
        let grid = self.as_standard_layout();
        let (mut idx, mut bad) = (0i64, -1i64);
        let st = unsafe {
            ffi::ndi_lower_index(
                <S::Elem as NdiElem>::DTYPE,
                grid.as_ptr() as *const c_void,
                grid.len() as i64,
                &x as *const S::Elem as *const c_void,
                1,
                &mut idx,
                &mut bad,
            )
        };
        if st == ffi::NDI_NAN_QUERY {
            unimplemented!("failed to convert {x:?} to usize")
        }
        assert!(st == ffi::NDI_OK, "ndi_lower_index: {}", ffi::last_error());
        idx as usize
    }
}
