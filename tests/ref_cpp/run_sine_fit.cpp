// Driver for the reference's own C++ test `test_cpp_sine_fit` (/root/reference/test/test_curve.cpp:31-93):
// fpCurve::new_fit (interpolating spline of sin on 201 points), ddx of orders 0..3 at 200 mid points,
// integral -- all through the handle C-ABI (fitpack_curve_c_*), here served by libfitpack_b200.so.
// The reference test source is compiled where it lies (tests/ref_cpp/build.py), never copied.
#include <cstdio>
extern "C" bool test_cpp_sine_fit();
int main()
{
    const bool ok = test_cpp_sine_fit();
    std::printf("test_cpp_sine_fit %s\n", ok ? "PASS" : "FAIL");
    return ok ? 0 : 1;
}
