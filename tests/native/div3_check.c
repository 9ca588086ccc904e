/* bc7_setup's div3 (fastc_b200/csrc/bc7.cu): q = RN(x * y), y = RN(1/3); RN(fma(fma(-q, 3, x), y, q)) must be the
 * IEEE quotient x / 3 (the reference divides the covariance sums by 3.0f, RGBAEndpoints.cpp:391, trap T8).
 * usage: div3_check FIRST_EXPONENT LAST_EXPONENT  -- checks every float with a biased exponent in that range
 * (1 254 = every positive normal float: ~40 s; the quotient is odd in x, so positive inputs suffice). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char **argv) {
  const uint32_t e0 = argc > 1 ? (uint32_t)atoi(argv[1]) : 1, e1 = argc > 2 ? (uint32_t)atoi(argv[2]) : 254;
  const float y = 0.3333333432674407958984375f;
  if (y != 1.0f / 3.0f) { puts("constant is not RN(1/3)"); return 2; }
  uint64_t bad = 0, n = 0;
  for (uint32_t bits = e0 << 23; bits < ((e1 + 1) << 23); bits++, n++) {
    float x;
    memcpy(&x, &bits, 4);
    const float q = x * y;
    const float res = fmaf(fmaf(-q, 3.0f, x), y, q);
    if (res != x / 3.0f && bad++ < 5) printf("mismatch x=%a ref=%a got=%a\n", x, x / 3.0f, res);
  }
  printf("checked %llu bad %llu\n", (unsigned long long)n, (unsigned long long)bad);
  return bad != 0;
}
