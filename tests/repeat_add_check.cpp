#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#define WRCU_HOSTEMU 1
#include "../webrender_b200/csrc/hostemu_shim.h"
#include "../webrender_b200/csrc/repeat_add.cuh"
static float naive(float x, float s, int n) { for (int i = 0; i < n; i++) { volatile float t = x + s; x = t; } return x; }
static uint32_t rng_state = 987654321;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5; return rng_state; }
static float rf(float lo, float hi) { return lo + (hi - lo) * (rnd() / 4294967296.0f); }
int main() {
  long bad = 0, total = 0;
  for (int it = 0; it < 1500000; it++) {
    float x, s; int n = rnd() % 4000;
    switch (it % 8) {
      case 0: x = rf(-10, 10); s = rf(-0.01f, 0.01f); break;
      case 1: x = rf(0, 1); s = rf(0, 0.002f); break;
      case 2: x = rf(-4000, 4000); s = rf(-3, 3); break;
      case 3: x = rf(0, 4000); s = (float)(rnd() % 64) / 64.0f; break;      // exact / tie-prone steps
      case 4: x = (float)(rnd() % 4096) / 8.0f; s = 0.5f / (1 << (rnd() % 12)); break;
      case 5: x = rf(-1e-3f, 1e-3f); s = rf(-1e-5f, 1e-5f); break;
      case 6: x = rf(100, 200); s = -rf(0, 1.0f); break;                    // crossing towards zero and sign change
      default: x = rf(-1, 1) * powf(2.0f, (float)(rnd() % 40) - 20); s = rf(-1, 1) * powf(2.0f, (float)(rnd() % 40) - 30); break;
    }
    float a = naive(x, s, n), b = wr_repeat_add(x, s, n);
    total++;
    if (memcmp(&a, &b, 4) != 0) { if (bad < 10) printf("MISMATCH x=%a s=%a n=%d naive=%a fast=%a\n", x, s, n, a, b); bad++; }
  }
  printf("total %ld bad %ld\n", total, bad);
  return bad != 0;
}
