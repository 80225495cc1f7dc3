#!/usr/bin/env python3
"""Transcribe the reference's own golden vectors into tests/golden/reference_vectors.json.

Run once in the build container (where /root/reference is mounted); the JSON it writes is
committed and is the only thing the tests read -- /root/reference does not exist on the GPU
box.  Expected-value arrays are pulled out of the Rust sources by line range so that no digit
is retyped by hand; the (small) inputs of each case are restated here next to the citation.

Nothing is *computed* here: the reference is Rust and cannot be executed in this image.
"""
import json
import os
import re
import sys

REF = os.environ.get("NDI_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")

NUM = re.compile(r"[-+]?(?:\d[\d_]*\.?[\d_]*(?:[eE][-+]?\d+)?|\.\d+)(?:f32|f64)?")


def lines(path, lo, hi):
    with open(os.path.join(REF, path)) as f:
        src = f.read().split("\n")
    # doc-tests live in `///` comments: drop the comment leader so the body parses like code
    return "\n".join(re.sub(r"^\s*///\s?", "", ln) for ln in src[lo - 1:hi])


def numbers(text):
    out = []
    for tok in NUM.findall(text):
        tok = tok.replace("_", "").replace("f32", "").replace("f64", "")
        out.append(float(tok))
    return out


def array_after(path, lo, hi, marker="let expect = array!["):
    """numbers of the array![...] literal that starts at `marker` inside [lo, hi]"""
    text = lines(path, lo, hi)
    start = text.index(marker) + len(marker)
    depth, i = 1, start
    while depth:
        c = text[i]
        depth += c == "["
        depth -= c == "]"
        i += 1
    body = re.sub(r"//[^\n]*", "", text[start:i - 1])
    return numbers(body)


DATA12 = [1.0, 2.0, 2.5, 2.5, 3.0, 2.0, 1.0, -2.0, 3.0, 5.0, 6.3, 8.0]
CS = "tests/cubic_spline_strat.rs"

cases = []


def cubic(name, rng, data, bc, extrapolate, query, x=None, dtype="f64", rel=1e-3, abs_=None, path=CS,
          expect=None, shape=None):
    lo, hi = rng
    if expect is None:
        expect = array_after(path, lo, hi)
    cases.append(dict(kind="cubic", name=name, ref=f"{path}:{lo}-{hi}", dtype=dtype, x=x, data=data,
                      bc=bc, extrapolate=extrapolate, query=query, expect=expect, expect_shape=shape,
                      tol=dict(rel=rel, abs=abs_ if abs_ is not None else
                               (1.1920929e-07 if dtype == "f32" else 2.220446049250313e-16))))


# ---- cubic spline -------------------------------------------------------------------------
cubic("doctest_wikipedia", (62, 82), [0.5, 0.0, 3.0], dict(kind="NotAKnot"), False,
      dict(linspace=[-1.0, 3.0, 10]), x=[-1.0, 0.0, 3.0], rel=0.0,
      path="src/interp1d/strategies/cubic_spline.rs",
      expect=array_after("src/interp1d/strategies/cubic_spline.rs", 62, 82, "let expect = array!["))
cubic("interp_natural", (10, 27), [1.0, 2.0, 3.0, 4.0, 3.0, 2.0, 1.0, 0.0, 2.0, 4.0, 6.0, 8.0],
      dict(kind="Natural"), False, dict(linspace=[0.0, 11.0, 30]))
cubic("extrapolate_natural", (58, 105), DATA12, dict(kind="Natural"), True, dict(linspace=[-3.0, 15.0, 30]))
cubic("extrapolate_not_a_knot", (108, 154), DATA12, dict(kind="NotAKnot"), True,
      dict(linspace=[-3.0, 15.0, 30]), dtype="f32",
      expect=array_after(CS, 108, 154, "let expect = array!["))
cubic("not_a_knot_3_values", (157, 188), [1.0, 2.0, 0.0], dict(kind="NotAKnot"), True,
      dict(linspace=[-1.0, 3.0, 15]))
cubic("multidim_multi_bounds", (191, 255), [[0.5, 1.0], [0.0, 1.5], [3.0, 0.5]],
      dict(kind="Individual", rows=[dict(kind="Natural"),
                                    dict(kind="Mixed", left=dict(kind="NotAKnot"),
                                         right=dict(kind="FirstDeriv", value=0.5))]),
      True, dict(linspace=[-2.0, 4.0, 15]), x=[-1.0, 0.0, 3.0], shape=[15, 2],
      expect=array_after(CS, 191, 255, "let expect = stack!["))
cubic("extrapolate_clamped", (258, 305), DATA12, dict(kind="Clamped"), True, dict(linspace=[-3.0, 15.0, 30]))
cubic("extrapolate_deriv1", (308, 358), DATA12,
      dict(kind="Individual", rows=[dict(kind="Mixed", left=dict(kind="FirstDeriv", value=-0.1),
                                         right=dict(kind="FirstDeriv", value=-0.5))]),
      True, dict(linspace=[-3.0, 15.0, 30]))
cubic("extrapolate_deriv2", (361, 411), DATA12,
      dict(kind="Individual", rows=[dict(kind="Mixed", left=dict(kind="SecondDeriv", value=-0.1),
                                         right=dict(kind="SecondDeriv", value=-0.5))]),
      True, dict(linspace=[-3.0, 15.0, 30]))
cubic("extrapolate_periodic", (455, 501), [1.0, 2.0, 2.5, 2.5, 3.0, 2.0, 1.0, -2.0, 3.0, 5.0, 6.3, 1.0],
      dict(kind="Periodic"), True, dict(linspace=[-3.0, 15.0, 30]))
cubic("extrapolate_periodic_multidim", (504, 537), [[0.5, 1.0], [0.0, 1.5], [0.0, 1.5], [0.5, 1.0]],
      dict(kind="Periodic"), True, dict(linspace=[-1.5, 3.5, 15]), x=[-1.0, 0.0, 2.0, 3.0], shape=[15, 2])
cubic("extrapolate_periodic_len3", (540, 573), [0.5, 0.0, 0.5], dict(kind="Periodic"), True,
      dict(linspace=[-1.5, 3.5, 15]), x=[-1.0, 0.0, 3.0])
cubic("extrapolate_periodic_len3_multidim", (576, 609), [[0.5, 1.0], [0.0, 2.5], [0.5, 1.0]],
      dict(kind="Periodic"), True, dict(linspace=[-1.5, 3.5, 15]), x=[-1.0, 0.0, 3.0], shape=[15, 2])

# the stack![Axis(1), ...] literal in multidim_multi_bounds starts with the tokens "Axis(1)":
# drop the stray "1" the number regex picks up from it.
for c in cases:
    if c["name"] == "multidim_multi_bounds":
        raw = array_after(CS, 191, 255, "let expect = stack![")
        assert raw[0] == 1.0 and len(raw) == 31, (raw[:3], len(raw))
        raw = raw[1:]
        c["expect"] = [raw[col * 15 + r] for r in range(15) for col in range(2)]

# ---- bilinear 11x11 matrix (tests/interp2d.rs:85-237) ---------------------------------------
cases.append(dict(kind="bilinear", name="interpolate_array", ref="tests/interp2d.rs:85-237", dtype="f64",
                  x=[1.0, 2.0, 3.0], y=[4.0, 5.0, 6.0], data=dict(linspace=[0.0, 8.0, 9], shape=[3, 3]),
                  qx=dict(linspace=[1.0, 3.0, 11], repeat_each=11),
                  qy=dict(linspace=[4.0, 6.0, 11], tile=11),
                  expect=array_after("tests/interp2d.rs", 85, 237), expect_shape=[11, 11],
                  tol=dict(abs=2.220446049250313e-16, rel=0.0)))

# ---- scalar / small exact cases, restated with their citations ---------------------------------
Y10 = [1.5, 2.0, 3.0, 4.0, 5.0, 7.0, 7.0, 8.0, 9.0, 10.5]
UPDOWN = [1.0, 2.0, 3.0, 4.0, 5.0, 5.0, 4.0, 3.0, 2.0, 1.0]
XM4 = [-4.0, -3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0, 4.0, 5.0]


def linear(name, ref, data, queries, expect, x=None, extrapolate=False, dtype="f64", abs_=0.0):
    cases.append(dict(kind="linear", name=name, ref=ref, dtype=dtype, x=x, data=data, extrapolate=extrapolate,
                      query=queries, expect=expect, tol=dict(abs=abs_, rel=0.0)))


EPS = 2.220446049250313e-16
linear("interp_y_only", "tests/interp1d.rs:21-30", Y10, [0.0, 9.0, 4.5, 0.25, 8.75], [1.5, 10.5, 6.0, 1.625, 10.125])
linear("extrapolate_y_only", "tests/interp1d.rs:33-40", [1.0, 2.0, 1.5], [-1.0, 3.0], [0.0, 1.0], extrapolate=True)
linear("interp_with_x_and_y", "tests/interp1d.rs:43-54", Y10, [-4.0, 5.0, 0.5, -3.75, 4.75],
       [1.5, 10.5, 6.0, 1.625, 10.125], x=XM4)
linear("interp_with_x_and_y_expspaced", "tests/interp1d.rs:57-69", UPDOWN, [1.0, 512.0, 42.0, 365.0],
       [1.0, 1.0, 4.6875, 1.57421875], x=[1.0, 2.0, 4.0, 8.0, 16.0, 32.0, 64.0, 128.0, 256.0, 512.0])
linear("extrapolate_with_x_and_y", "tests/interp1d.rs:72-80", [1.0, 0.0, 1.5], [-1.0, 2.0], [2.0, 3.0],
       x=[0.0, 1.0, 1.5], extrapolate=True)
linear("interp_array", "tests/interp1d.rs:83-90", UPDOWN, [[1.0, 2.0, 9.0], [4.0, 5.0, 7.5]],
       [[2.0, 3.0, 1.0], [5.0, 5.0, 2.5]])
linear("interp_view_array", "tests/interp1d.rs:143-155", [10.0, 9.0, 8.0, 7.0, 6.0, 5.0, 4.0, 3.0, 2.0, 1.0],
       [-4.0, 5.0, 0.0, -3.5, 4.75], [10.0, 1.0, 6.0, 9.5, 1.25], x=XM4)
linear("interp_multi_fn_single", "tests/interp1d.rs:158-174",
       [[0.1, 0.2, 0.3, 0.4, 0.5], [2.0, 2.0, 3.0, 4.0, 5.0], [10.0, 20.0, 30.0, 40.0, 50.0],
        [20.0, 40.0, 60.0, 80.0, 100.0]], [1.5], [[1.05, 1.1, 1.65, 2.2, 2.75]], x=[1.0, 2.0, 3.0, 4.0], abs_=EPS)
linear("interp_multi_fn_array", "tests/interp1d.rs:175-195",
       [[0.1, 0.2, 0.3, 0.4, 0.5], [2.0, 2.0, 3.0, 4.0, 5.0], [10.0, 20.0, 30.0, 40.0, 50.0],
        [20.0, 40.0, 60.0, 80.0, 100.0]], [[1.0, 1.5], [3.5, 4.0]],
       [[[0.1, 0.2, 0.3, 0.4, 0.5], [1.05, 1.1, 1.65, 2.2, 2.75]],
        [[15.0, 30.0, 45.0, 60.0, 75.0], [20.0, 40.0, 60.0, 80.0, 100.0]]], x=[1.0, 2.0, 3.0, 4.0], abs_=EPS)
linear("lib_doc_1d", "src/lib.rs:41-47", [0.0, 1.0, 1.5, 1.0, 0.0], [3.5, 0.0, 0.5, 1.5], [0.5, 0.0, 0.5, 1.25])
linear("lib_doc_1d_multidim", "src/lib.rs:55-71", [[0.0, 1.0], [1.0, 2.0], [1.5, 2.5], [1.0, 2.0]],
       [0.5, 4.0], [[-0.5, 0.5], [1.0, 2.0]], x=[1.0, 2.0, 3.0, 4.0], extrapolate=True)
linear("interp1d_doc_scalar", "src/interp1d/mod.rs:99-106", [1.0, 1.5, 2.0], [1.5], [1.25], x=[1.0, 2.0, 3.0])
linear("interp1d_doc_interp", "src/interp1d/mod.rs:135-145", [[0.0, 2.0, 4.0], [0.5, 2.5, 3.5], [1.0, 3.0, 3.0]],
       [0.5], [[0.25, 2.25, 3.75]], abs_=EPS)
linear("interp1d_doc_array", "src/interp1d/mod.rs:185-195", [0.0, 0.5, 1.0], [0.5, 1.0, 1.5], [0.25, 0.5, 0.75],
       x=[0.0, 1.0, 2.0], abs_=EPS)
linear("interp1d_doc_array_into", "src/interp1d/mod.rs:234-267", [[0.0, 2.0], [0.5, 2.5], [1.0, 3.0]],
       [[0.0, 0.5], [1.0, 1.5]], [[[0.0, 2.0], [0.25, 2.25]], [[0.5, 2.5], [0.75, 2.75]]], x=[0.0, 1.0, 2.0], abs_=EPS)


def bil(name, ref, data, qx, qy, expect, x=None, y=None, dtype="f64", abs_=0.0, extrapolate=False):
    cases.append(dict(kind="bilinear", name=name, ref=ref, dtype=dtype, x=x, y=y, data=data, qx=qx, qy=qy,
                      extrapolate=extrapolate, expect=expect, tol=dict(abs=abs_, rel=0.0)))


I34 = [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]
F34 = [[float(v) for v in r] for r in I34]
bil("cornerns_only_data_no_axis", "tests/interp2d.rs:27-34", I34, [0, 2, 2, 0], [0, 3, 0, 3], [1, 12, 9, 4], dtype="i32")
bil("cornerns_only_x_axis", "tests/interp2d.rs:37-47", I34, [1, 3, 3, 1], [0, 3, 0, 3], [1, 12, 9, 4], x=[1, 2, 3], dtype="i32")
bil("cornerns_only_y_axis", "tests/interp2d.rs:50-60", F34, [0.0, 2.0, 2.0, 0.0], [-3.0, 0.0, -3.0, 0.0],
    [1.0, 12.0, 9.0, 4.0], y=[-3.0, -2.0, -1.0, 0.0])
ND = [[[[1.0, 10.0], [-1.0, -10.0]], [[2.0, 20.0], [-2.0, -20.0]]],
      [[[3.0, 30.0], [-3.0, -30.0]], [[5.0, 50.0], [-5.0, -50.0]]]]
bil("interp_nd_data_single", "tests/interp2d.rs:241-258", ND, [0.0], [0.5], [[[1.5, 15.0], [-1.5, -15.0]]], abs_=EPS)
bil("interp_nd_data_array", "tests/interp2d.rs:260-264", ND, [0.0, 0.5], [0.5, 1.0],
    [[[1.5, 15.0], [-1.5, -15.0]], [[3.5, 35.0], [-3.5, -35.0]]], abs_=EPS)
bil("lib_doc_2d", "src/lib.rs:79-88", [[1.0, 2.0, 2.5], [3.0, 4.0, 3.5]], [0.0, 0.0, 1.0], [0.5, 0.5, 2.0], [1.5, 1.5, 3.5])
bil("lib_doc_2d_multidim", "src/lib.rs:96-114",
    [[[1.0, -1.0], [2.0, -2.0], [3.0, -3.0]], [[4.0, -4.0], [5.0, -5.0], [6.0, -6.0]],
     [[7.0, -7.0], [8.0, -8.0], [9.0, -9.0]], [[7.5, -7.5], [8.5, -8.5], [9.5, -9.5]]],
    [1.5, 1.5, 1.5], [2.0, 2.0, 2.5], [[3.5, -3.5], [3.5, -3.5], [4.0, -4.0]],
    x=[1.0, 2.0, 3.0, 4.0], y=[1.0, 2.0, 3.0])
bil("interp2d_doc_scalar", "src/interp2d/mod.rs:96-105", [[1.0, 2.0], [3.0, 4.0]], [0.0], [0.5], [1.5])

# ---- out-of-bounds pins -------------------------------------------------------------------------
oob = [
    dict(kind="oob1d", name="interp_y_only_out_of_bounds", ref="tests/interp1d.rs:93-103", data=[1.0, 2.0, 3.0],
         x=None, queries=[-0.1, 9.0]),
    dict(kind="oob1d", name="interp_with_x_and_y_out_of_bounds", ref="tests/interp1d.rs:106-120",
         data=[1.0, 2.0, 3.0], x=[-4.0, -3.0, 2.0], queries=[-4.1, 2.1]),
    dict(kind="oob1d_cubic", name="extrapolate_false", ref="tests/cubic_spline_strat.rs:46-55",
         data=[1.0, 2.0, 1.0], x=None, queries=[-0.5, 3.5]),
    dict(kind="oob2d", name="extrapolate", ref="tests/interp2d.rs:63-82", data=I34, dtype="i32",
         queries=[[-1, 1, 0], [1, -1, 1], [3, 1, 0], [1, 4, 1]]),  # (x, y, failing axis)
]
cases += oob

# ---- lower-index and monotonic pins (src/vector_extensions.rs unit tests) -------------------------
index_cases = [
    dict(kind="index", name="linspace_0_10_11", ref="src/vector_extensions.rs:221-265", grid="linspace",
         pairs=[[0, -1.0], [9, 25.0], [0, 0.0], [9, 10.0]] + [[i, float(i)] for i in range(10)]
         + [[i // 10, i / 10.0] for i in range(100)] + [[9, "inf"], [0, "-inf"]]),
    dict(kind="index", name="exp2", ref="src/vector_extensions.rs:273-295", grid="exp2",
         pairs=[[i, float(2 ** i)] for i in range(10)] + [[i // 10, {"exp2": i / 10.0}] for i in range(100)]
         + [[9, 1024.0], [0, 1.0]]),
    dict(kind="index", name="ln_1p", ref="src/vector_extensions.rs:297-302", grid="ln_1p",
         pairs=[[i // 10, {"ln_1p": i / 10.0}] for i in range(100)]),
    dict(kind="index_nan", name="test_nan", ref="src/vector_extensions.rs:267-271", grid="linspace"),
]
cases += index_cases
mono = [
    ("f64", [1.1, 2.0, 3.123, 4.5], "RisingStrict", ":319-322"),
    ("f64", [1.1, 2.0, 3.123, 3.123, 4.5], "Rising", ":325-328"),
    ("f64", [5.8, 4.123, 3.1, 2.0, 1.0], "FallingStrict", ":331-334"),
    ("f64", [5.8, 4.123, 3.1, 3.1, 2.0, 1.0], "Falling", ":337-340"),
    ("f64", [1.1, 2.0, 3.123, 3.120, 4.5], "NotMonotonic", ":343-346"),
    ("i32", [1, 2, 3, 4, 5], "RisingStrict", ":350-353"),
    ("i32", [1, 2, 3, 3, 4, 5], "Rising", ":356-359"),
    ("i32", [5, 4, 3, 2, 1], "FallingStrict", ":362-365"),
    ("i32", [5, 4, 3, 3, 2, 1], "Falling", ":368-371"),
    ("i32", [1, 2, 3, 2, 4, 5], "NotMonotonic", ":374-377"),
    ("i32", [1, 1, 2, 3, 4, 5], "Rising", ":387-390"),
    ("i32", [1, 1, 1], "NotMonotonic", ":393-396"),
    ("i32", [1], "NotMonotonic", ":399-402"),
]
for dt, arr, exp, where in mono:
    cases.append(dict(kind="monotonic", name=f"mono_{dt}_{exp}_{len(arr)}", ref="src/vector_extensions.rs" + where,
                      dtype=dt, x=arr, stride=1, expect=exp))
cases.append(dict(kind="monotonic", name="test_ordered_view_on_unordred_array", ref="src/vector_extensions.rs:380-384",
                  dtype="i32", x=[5, 4, 3, 2, 1], stride=-1, expect="RisingStrict"))

# ---- builder-error pins ----------------------------------------------------------------------------
cases.append(dict(kind="builder_errors", name="builder_errors", ref="tests/interp1d.rs:123-140; tests/interp2d.rs:280-329; "
                  "tests/cubic_spline_strat.rs:30-43,414-452",
                  note="restated directly in tests/test_reference_*.py (variant checks, no numeric vectors)"))

if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"reference not mounted at {REF}; the committed JSON is the artefact")
    with open(OUT, "w") as f:
        json.dump(dict(source="jonasBoss/ndarray-interp v0.6.0 test suite (transcribed, not computed)",
                       cases=cases), f, indent=1)
    print(f"wrote {len(cases)} cases to {OUT}")
