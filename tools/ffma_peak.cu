// standalone: prints the measured FFMA / FFMA2 peaks (see csrc/micro.cuh)
#include <cstdio>
#include "../s2s-ismr-unet_b200/csrc/micro.cuh"
int main() {
    float a = 0, b = 0;
    if (s2s::ffma_peak_measure(0, &a, 0) || s2s::ffma_peak_measure(1, &b, 0)) { printf("error: %s\n", s2s::last_error_ref().c_str()); return 1; }
    printf("ffma_peak: scalar FFMA %.2f TFLOP/s, packed FFMA2 %.2f TFLOP/s\n", a, b);
    return 0;
}
