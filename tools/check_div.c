// Exhaustive check that the 3-instruction Markstein sequence used by jss_div() (jss_device.cuh)
// equals IEEE fp32 division for integer x in [0, xmax] over integer y.  usage: check_div y xmax [y xmax ...]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
// exhaustive check: Markstein sequence == correctly rounded fp32 division for integer x in [0, xmax], integer y
int main(int argc, char **argv) {
    long bad = 0, n = 0;
    for (int a = 1; a + 1 < argc; a += 2) {
        int y = atoi(argv[a]), xmax = atoi(argv[a + 1]);
        volatile float fy = (float)y;
        volatile float ry = 1.0f / fy;
        for (int x = 0; x <= xmax; x++) {
            float fx = (float)x;
            volatile float q = fx * ry;
            float r = fmaf(-q, fy, fx);
            float q1 = fmaf(r, ry, q);
            float ref = fx / fy;
            n++;
            if (q1 != ref) { bad++; if (bad < 10) printf("mismatch x=%d y=%d %a %a\n", x, y, q1, ref); }
        }
    }
    printf("checked %ld pairs, %ld mismatches\n", n, bad);
    return bad != 0;
}
