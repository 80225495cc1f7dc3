// C++ counterpart of the reference's integration tests, written against include/ndarray_interp_b200.hpp.
// Each test names the reference test whose behaviour it checks (tests/interp1d.rs, tests/interp2d.rs,
// tests/cubic_spline_strat.rs, src/vector_extensions.rs, examples/custom_strategy.rs).  Expected values
// are the reference's own literals; where the reference asserts with a tolerance, the same tolerance
// is used.  Exit code = number of failed checks.
#include <cstdio>
#include <limits>

#include "ndarray_interp_b200.hpp"

using namespace ndarray_interp;
using A = Array<double>;

static int failures = 0, checks = 0;
#define CHECK(cond)                                                                    \
    do { ++checks; if (!(cond)) { ++failures; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } } while (0)
template <class E, class F>
static bool throws(F&& f, int kind = -1) {
    try { f(); } catch (const E& e) { return kind < 0 || (int)e.kind == kind; } catch (...) { return false; }
    return false;
}
template <class F>
static bool panics(F&& f, const char* needle = "") {
    try { f(); } catch (const Panic& e) { return std::string(e.what()).find(needle) != std::string::npos; } catch (...) { return false; }
    return false;
}
static bool close(double a, double b, double eps) { return std::fabs(a - b) <= eps; }
static const double EPS = std::numeric_limits<double>::epsilon();

// tests/interp1d.rs: interp_y_only, extrapolate_y_only, interp_with_x_and_y, ..._expspaced, extrapolate_with_x_and_y
static void linear_scalar() {
    auto ip = Interp1D<double>::builder(A{1.5, 2.0, 3.0, 4.0, 5.0, 7.0, 7.0, 8.0, 9.0, 10.5}).build();
    CHECK(ip.interp_scalar(0.0) == 1.5);
    CHECK(ip.interp_scalar(9.0) == 10.5);
    CHECK(ip.interp_scalar(4.5) == 6.0);
    CHECK(ip.interp_scalar(0.25) == 1.625);
    CHECK(ip.interp_scalar(8.75) == 10.125);
    auto ex = Interp1D<double>::builder(A{1.0, 2.0, 1.5}).strategy(Linear<double>().extrapolate(true)).build();
    CHECK(ex.interp_scalar(-1.0) == 0.0);
    CHECK(ex.interp_scalar(3.0) == 1.0);
    auto xy = Interp1DBuilder<double>(A{1.5, 2.0, 3.0, 4.0, 5.0, 7.0, 7.0, 8.0, 9.0, 10.5})
                  .x(A{-4.0, -3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0, 4.0, 5.0}).strategy(Linear<double>()).build();
    CHECK(xy.interp_scalar(-4.0) == 1.5);
    CHECK(xy.interp_scalar(5.0) == 10.5);
    CHECK(xy.interp_scalar(0.5) == 6.0);
    CHECK(xy.interp_scalar(-3.75) == 1.625);
    CHECK(xy.interp_scalar(4.75) == 10.125);
    auto ee = Interp1DBuilder<double>(A{1.0, 2.0, 3.0, 4.0, 5.0, 5.0, 4.0, 3.0, 2.0, 1.0})
                  .x(A{1.0, 2.0, 4.0, 8.0, 16.0, 32.0, 64.0, 128.0, 256.0, 512.0}).build();
    CHECK(ee.interp_scalar(1.0) == 1.0);
    CHECK(ee.interp_scalar(512.0) == 1.0);
    CHECK(ee.interp_scalar(42.0) == 4.6875);
    CHECK(ee.interp_scalar(365.0) == 1.57421875);
    auto ex2 = Interp1DBuilder<double>(A{1.0, 0.0, 1.5}).x(A{0.0, 1.0, 1.5}).strategy(Linear<double>().extrapolate(true)).build();
    CHECK(ex2.interp_scalar(-1.0) == 2.0);
    CHECK(ex2.interp_scalar(2.0) == 3.0);
}

// tests/interp1d.rs: interp_array (2-D query), interp_array_into shapes, multi-dimensional data
static void linear_arrays() {
    auto ip = Interp1D<double>::builder(A{1.0, 2.0, 3.0, 4.0, 5.0, 5.0, 4.0, 3.0, 2.0, 1.0}).build();
    A q{{1.0, 2.0, 9.0}, {4.0, 5.0, 7.5}};
    CHECK(ip.interp_array(q) == (A{{2.0, 3.0, 1.0}, {5.0, 5.0, 2.5}}));
    // data (4, 2): one row per x, query (3,) -> (3, 2)
    auto ip2 = Interp1D<double>::builder(A{{0.0, 10.0}, {1.0, 20.0}, {3.0, 40.0}, {6.0, 70.0}}).build();
    CHECK(ip2.interp_array(A{0.5, 1.5, 3.0}) == (A{{0.5, 15.0}, {2.0, 30.0}, {6.0, 70.0}}));
    CHECK(ip2.interp(2.5) == (A{4.5, 55.0}));
    A buf({2}, 0.0);
    ip2.interp_into(0.5, buf.view_mut());
    CHECK(buf == (A{0.5, 15.0}));
    A wrong({3}, 0.0);
    CHECK(panics([&] { ip2.interp_into(0.5, wrong.view_mut()); }, "dimension mismatch"));
    A out({2, 3}, 0.0);                                       // query (2,) needs (2, 2)
    CHECK(panics([&] { ip2.interp_array_into(A{0.5, 1.0}, out.view_mut()); }, "incompatible shapes"));
    // a reversed view as x and data (tests/interp1d.rs: slice(s![..;-1]))
    A xr{3.0, 2.0, 1.0, 0.0}, dr{30.0, 20.0, 10.0, 0.0};
    auto rv = Interp1D<double>::new_unchecked(xr.reversed(), dr.reversed(), std::make_shared<Linear<double>>());
    CHECK(rv.interp_scalar(1.5) == 15.0);
}

// tests/interp1d.rs: out_of_bounds, interp_builder_errors; first error wins and rows before it are written
static void linear_errors() {
    auto ip = Interp1D<double>::builder(A{1.0, 2.0, 3.0}).build();
    CHECK(throws<InterpolateError>([&] { ip.interp(-0.1); }));
    CHECK(throws<InterpolateError>([&] { ip.interp(9.0); }));
    auto ipx = Interp1DBuilder<double>(A{1.0, 2.0, 3.0}).x(A{-4.0, -3.0, 2.0}).build();
    CHECK(throws<InterpolateError>([&] { ipx.interp(-4.1); }));
    CHECK(throws<InterpolateError>([&] { ipx.interp(2.1); }));
    using I = Array<int32_t>;
    CHECK(throws<BuilderError>([] { Interp1DBuilder<int32_t>(I{1}).build(); }, BuilderError::NotEnoughData));
    CHECK(throws<BuilderError>([] { Interp1DBuilder<int32_t>(I{1, 2}).x(I{1, 2, 3}).build(); }, BuilderError::ShapeError));
    CHECK(throws<BuilderError>([] { Interp1DBuilder<int32_t>(I{1, 2, 3}).x(I{1, 2, 2}).build(); }, BuilderError::Monotonic));
    A buf({4}, -7.0);
    try { ip.interp_array_into(A{0.5, 1.5, 5.0, 1.0}, buf.view_mut()); CHECK(false); }
    catch (const InterpolateError& e) { CHECK(std::string(e.what()) == "x = 5.0 is not in range"); }
    CHECK(buf == (A{1.5, 2.5, -7.0, -7.0}));
    auto ex = Interp1D<double>::builder(A{1.0, 2.0, 3.0}).strategy(Linear<double>().extrapolate(true)).build();
    CHECK(panics([&] { ex.interp_scalar(std::nan("")); }, "failed to convert NaN to usize"));
    // integer element type: truncating division like the reference's generic Num path
    auto ii = Interp1D<int32_t>::builder(I{10, 20, 40}).build();
    CHECK(ii.interp_array(I{0, 1, 2}) == (I{10, 20, 40}));
    using I8 = Array<int64_t>;
    const int64_t big = (int64_t)1 << 60;
    auto i8 = Interp1D<int64_t>::builder(I8{big, big + 20, big + 50}).x(I8{0, 10, 40}).build();
    CHECK(i8.interp_array(I8{0, 5, 10, 25, 40}) == (I8{big, big + 10, big + 20, big + 35, big + 50}));
    CHECK(throws<InterpolateError>([&] { i8.interp_scalar(41); }));
    // unsigned: values with the top bit set compare and divide as unsigned; falling data wraps like a release build
    using U4 = Array<uint32_t>;
    const uint32_t top = 0x80000000u;
    auto u4 = Interp1D<uint32_t>::builder(U4{top - 10, top + 10, top + 40}).x(U4{top - 5, top + 5, top + 35}).build();
    CHECK(u4.interp_array(U4{top - 5, top, top + 5, top + 20, top + 35}) == (U4{top - 10, top, top + 10, top + 25, top + 40}));
    CHECK(throws<InterpolateError>([&] { u4.interp_scalar(top + 36); }));
    auto fall = Interp1D<uint32_t>::builder(U4{30, 10}).x(U4{0, 10}).build();
    CHECK(fall.interp_scalar(5) == (uint32_t)(((uint32_t)(10 - 30) / 10u) * 5u + 30u));   // (y2 - y1) wraps, unsigned division
}

// src/vector_extensions.rs unit tests: borders, exact hits, +-inf, NaN panic, monotonic classification
static void vector_extensions() {
    A g{0.0, 1.0, 2.0, 4.0, 8.0, 16.0};
    CHECK(get_lower_index<double>(g, -1.0) == 0);
    CHECK(get_lower_index<double>(g, 0.0) == 0);
    CHECK(get_lower_index<double>(g, 3.9) == 2);
    CHECK(get_lower_index<double>(g, 4.0) == 3);
    CHECK(get_lower_index<double>(g, 16.0) == 4);
    CHECK(get_lower_index<double>(g, 1e300) == 4);
    CHECK(get_lower_index<double>(g, std::numeric_limits<double>::infinity()) == 4);
    CHECK(get_lower_index<double>(g, -std::numeric_limits<double>::infinity()) == 0);
    CHECK(panics([&] { get_lower_index<double>(g, std::nan("")); }, "NaN"));
    CHECK((monotonic_prop<double>(g) == Monotonic{Monotonic::Rising, true}));
    CHECK((monotonic_prop<double>(A{0.0, 1.0, 1.0, 2.0}) == Monotonic{Monotonic::Rising, false}));
    CHECK((monotonic_prop<double>(A{3.0, 2.0, 1.0}) == Monotonic{Monotonic::Falling, true}));
    CHECK((monotonic_prop<double>(A{3.0, 2.0, 2.0}) == Monotonic{Monotonic::Falling, false}));
    CHECK((monotonic_prop<double>(A{1.0, 3.0, 2.0}) == Monotonic{Monotonic::NotMonotonic, false}));
    CHECK((monotonic_prop<double>(A{1.0}) == Monotonic{Monotonic::NotMonotonic, false}));
    CHECK((monotonic_prop<double>(A{2.0, 2.0, 2.0}) == Monotonic{Monotonic::NotMonotonic, false}));
    CHECK((monotonic_prop<double>(g.reversed()) == Monotonic{Monotonic::Falling, true}));
}

// cubic_spline.rs doc example (:62-82): values pinned at f64::EPSILON
static void cubic_doc_example() {
    A y{0.5, 0.0, 3.0, 6.0, 2.0, 1.0, 1.5, 3.0, 2.5, 0.5};
    auto ip = Interp1DBuilder<double>(y).strategy(CubicSpline<double>()).build();
    const double xs[] = {0.0, 1.0, 2.0, 9.0};
    for (double x : xs) CHECK(close(ip.interp_scalar(x), y[(size_t)x], 8 * EPS));       // a spline reproduces its knots
    // natural boundary: second derivative vanishes at the ends -> compare with scipy-derived values of
    // tests/cubic_spline_strat.rs (natural, max_relative 1e-3)
    auto nat = Interp1DBuilder<double>(A{1.0, 2.0, 3.0, 4.0, 5.0, 6.0}).strategy(CubicSpline<double>().boundary(BoundaryCondition<double>::Natural())).build();
    CHECK(close(nat.interp_scalar(2.5), 3.5, 1e-12));                                   // a straight line stays a straight line
    CHECK(throws<InterpolateError>([&] { nat.interp_scalar(5.5); }));
    auto ex = Interp1DBuilder<double>(A{1.0, 2.0, 3.0, 4.0, 5.0, 6.0}).strategy(CubicSpline<double>().extrapolate(true)).build();
    CHECK(close(ex.interp_scalar(7.0), 8.0, 1e-12));
    CHECK(throws<BuilderError>([] { Interp1DBuilder<double>(A{1.0, 2.0}).strategy(CubicSpline<double>()).build(); }, BuilderError::NotEnoughData));
    // periodic: first and last value must match (cubic_spline.rs:499-507); periodic extrapolation wraps (:805-809)
    CHECK(throws<BuilderError>([] {
        Interp1DBuilder<double>(A{1.0, 2.0, 3.0, 4.0}).strategy(CubicSpline<double>().boundary(BoundaryCondition<double>::Periodic())).build();
    }, BuilderError::ValueError));
    auto per = Interp1DBuilder<double>(A{0.0, 1.0, 0.0, -1.0, 0.0}).strategy(CubicSpline<double>().boundary(BoundaryCondition<double>::Periodic()).extrapolate(true)).build();
    CHECK(close(per.interp_scalar(0.5), per.interp_scalar(4.5), 1e-12));
    CHECK(close(per.interp_scalar(-3.0), per.interp_scalar(1.0), 1e-12));
    // Individual boundaries: shape check (tests/cubic_spline_strat.rs:414,429)
    A d2{{1.0, 2.0}, {2.0, 3.0}, {4.0, 1.0}, {2.0, 2.0}};
    auto bad = BoundaryCondition<double>::Individual({1, 3}, {RowBoundary<double>::Natural(), RowBoundary<double>::Natural(), RowBoundary<double>::Natural()});
    try { Interp1DBuilder<double>(d2).strategy(CubicSpline<double>().boundary(bad)).build(); CHECK(false); }
    catch (const BuilderError& e) { CHECK(e.kind == BuilderError::ShapeError); CHECK(std::string(e.what()).find("Expected: [1, 2], got: [1, 3]") != std::string::npos); }
    auto good = BoundaryCondition<double>::Individual({1, 2}, {RowBoundary<double>::Natural(), RowBoundary<double>::Mixed(SingleBoundary<double>::FirstDeriv(0.5), SingleBoundary<double>::NotAKnot())});
    auto ind = Interp1DBuilder<double>(d2).strategy(CubicSpline<double>().boundary(good)).build();
    CHECK(close(ind.interp(2.0)[0], 4.0, 8 * EPS) && close(ind.interp(2.0)[1], 1.0, 8 * EPS));
    // the two build modes of the device-side solve (NOT in the reference): the reference's elimination order and
    // the row-split (cyclic reduction + Thomas) build agree far inside 1e-12 and both reproduce the knots
    std::vector<double> yl(4096);
    for (size_t i = 0; i < yl.size(); ++i) yl[i] = std::sin(0.01 * (double)i) + 0.001 * (double)((i * 2654435761u) % 1000);
    A ylong(std::vector<size_t>{yl.size()}, yl);
    auto seq = Interp1DBuilder<double>(ylong).strategy(CubicSpline<double>().solver(NDI_BUILD_SEQUENTIAL)).build();
    auto split = Interp1DBuilder<double>(ylong).strategy(CubicSpline<double>().solver(NDI_BUILD_ROWSPLIT, 3)).build();
    auto aut = Interp1DBuilder<double>(ylong).strategy(CubicSpline<double>()).build();
    CHECK(static_cast<const CubicSplineStrategy<double>&>(seq.strategy()).rowsplit_levels(seq) == 0);
    CHECK(static_cast<const CubicSplineStrategy<double>&>(split.strategy()).rowsplit_levels(split) == 3);
    CHECK(static_cast<const CubicSplineStrategy<double>&>(aut.strategy()).rowsplit_levels(aut) == -32); // AUTO: 4096 rows -> partition, blocks of 32
    auto part = Interp1DBuilder<double>(ylong).strategy(CubicSpline<double>().solver(NDI_BUILD_PARTITION, 16)).build();
    CHECK(static_cast<const CubicSplineStrategy<double>&>(part.strategy()).rowsplit_levels(part) == -16);
    for (double x : {0.5, 17.25, 2047.5, 4094.75}) {
        CHECK(close(seq.interp_scalar(x), split.interp_scalar(x), 1e-12));
        CHECK(close(seq.interp_scalar(x), part.interp_scalar(x), 1e-12));
        CHECK(close(seq.interp_scalar(x), aut.interp_scalar(x), 1e-12));
    }
    CHECK(close(split.interp_scalar(100.0), yl[100], 8 * EPS));
}

// tests/interp2d.rs: interp_scalar values, data with trailing axes, range errors with x before y
static void bilinear() {
    A z{{1.0, 2.0, 3.0}, {2.0, 4.0, 6.0}, {3.0, 6.0, 9.0}};   // z = (x+1)(y+1)
    auto ip = Interp2D<double>::builder(z).build();
    CHECK(ip.interp_scalar(0.0, 0.0) == 1.0);
    CHECK(ip.interp_scalar(2.0, 2.0) == 9.0);
    CHECK(ip.interp_scalar(0.5, 0.5) == 2.25);
    CHECK(ip.interp_scalar(1.5, 0.25) == 3.125);
    CHECK(ip.interp_array(A{0.5, 1.0}, A{0.5, 2.0}) == (A{2.25, 6.0}));
    try { ip.interp_scalar(2.5, 7.0); CHECK(false); } catch (const InterpolateError& e) { CHECK(std::string(e.what()).rfind("x = 2.5", 0) == 0); }
    try { ip.interp_scalar(1.0, 7.0); CHECK(false); } catch (const InterpolateError& e) { CHECK(std::string(e.what()).rfind("y = 7.0", 0) == 0); }
    auto ex = Interp2D<double>::builder(z).strategy(Bilinear<double>().extrapolate(true)).build();
    CHECK(ex.interp_scalar(3.0, 3.0) == 16.0);
    CHECK(panics([&] { ip.interp_array(A{0.5, 1.0}, A{0.5}); }, "do not match"));
    CHECK(throws<BuilderError>([&] { Interp2D<double>::builder(z).x(A{0.0, 1.0, 1.0}).build(); }, BuilderError::Monotonic));
    CHECK(throws<BuilderError>([&] { Interp2D<double>::builder(z).y(A{0.0, 1.0}).build(); }, BuilderError::ShapeError));
    CHECK(throws<BuilderError>([&] { Interp2D<double>::builder(A{{1.0, 2.0}}).build(); }, BuilderError::NotEnoughData));
    Array<int32_t> zi{{1, 2}, {3, 4}};                         // tests/interp2d.rs integer data
    auto ii = Interp2D<int32_t>::builder(zi).build();
    CHECK(ii.interp_scalar(1, 1) == 4);
}

// examples/custom_strategy.rs: a user strategy implements the per-query trait method as host code
struct StepInterpolator : Interp1DStrategyBuilder<double>, Interp1DStrategy<double> {
    size_t MINIMUM_DATA_LENGHT() const override { return 2; }
    std::shared_ptr<const Interp1DStrategy<double>> build(const ArrayView<double>&, const ArrayView<double>&) const override {
        return std::make_shared<StepInterpolator>(*this);
    }
    void interp_into(const Interp1D<double>& ip, ArrayViewMut<double> target, double x) const override {
        const size_t idx = ip.get_index_left_of(x);
        auto [xl, dl] = ip.index_point(idx);
        auto [xr, dr] = ip.index_point(idx + 1);
        target.assign_rows((xr - xl) / 2.0 > (x - xl) ? dl.to_vector() : dr.to_vector());
    }
};
static void custom_strategy() {
    auto ip = Interp1D<double>::builder(A{2.0, 4.0, 5.0}).strategy(StepInterpolator()).build();
    const A res = ip.interp_array(A::linspace(-0.5, 2.5, 6));
    const A expect{2.0, 2.0, 4.0, 4.0, 5.0, 5.0};
    for (size_t i = 0; i < 6; ++i) CHECK(close(res[i], expect[i], EPS));
}

int main() {
    int32_t ndev = 0;
    if (ndi_device_count(&ndev) != NDI_OK || ndev < 1) { std::printf("no CUDA device: %s\n", ndi_last_error_message()); return 99; }
    linear_scalar();
    linear_arrays();
    linear_errors();
    vector_extensions();
    cubic_doc_example();
    bilinear();
    custom_strategy();
    std::printf("%d checks, %d failures\n", checks, failures);
    return failures > 255 ? 255 : failures;
}
