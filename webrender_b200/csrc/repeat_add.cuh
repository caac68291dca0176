// repeat_add.cuh — n sequential fp32 additions x += s, bit for bit, in
// O(binades crossed) instead of O(n).
//
// The reference walks edges and spans with running sums (Edge::nextRow,
// rasterize.h:880-884; interp += interp_step per chunk), so row 2000 of a tall
// primitive holds the result of 2000 rounded additions.  A tile CTA needs that
// value without walking from row 0.  Inside one binade every float is a
// multiple of the same ulp u, so RN(x + s) = x + d with a CONSTANT increment
// d (s rounded to a multiple of u) — unless x + s falls exactly half way
// between two floats (a tie, resolved by mantissa parity).  Hence:
//   take one real step to learn d; if it was not a tie and stayed in the binade,
//   jump m steps at once (x + m*d is exactly representable: one fused
//   multiply-add), m = steps left before the binade boundary; repeat.
// Ties and tiny/denormal/non-finite values fall back to real steps.
#pragma once
#include "wrcu_internal.h"

WRD_SHARED float wr_repeat_add(float x, float s, int n) {
  if (n <= 0) return x;
  if (s == 0.0f || !(fabsf(x) < 3.0e38f) || !(fabsf(s) < 3.0e38f)) {
    for (int i = 0; i < n && i < 4; i++) x = __fadd_rn(x, s);  // inf/nan/zero step: a few steps settle it
    return x;
  }
  while (n > 0) {
    const float y = __fadd_rn(x, s);
    n--;
    if (n == 0) return y;
    const uint32_t bx = __float_as_uint(x), by = __float_as_uint(y);
    const int ex = (int)((bx >> 23) & 0xFF), ey = (int)((by >> 23) & 0xFF);
    // jump only inside a normal binade, with the sign kept
    if (ex == ey && ex > 24 && ex < 254 && ((bx ^ by) >> 31) == 0) {
      const float d = __fsub_rn(y, x);  // exact: both are multiples of the binade's ulp
      // tie test: exact residual of the addition (TwoSum); a tie leaves exactly ulp/2
      const float bb = __fsub_rn(y, x);
      const float err = __fadd_rn(__fsub_rn(x, __fsub_rn(y, bb)), __fsub_rn(s, bb));
      const float ulp = __uint_as_float((uint32_t)(ex - 23) << 23);
      if (fabsf(err) != 0.5f * ulp && d != 0.0f) {
        // integer mantissa arithmetic: X in [2^23, 2^24), D = d / ulp
        const int32_t X = (int32_t)((by & 0x007FFFFFu) | 0x00800000u);
        const float dq = __fdiv_rn(d, ulp);  // exact power-of-two scaling
        const int32_t D = (int32_t)dq * ((by >> 31) ? -1 : 1);  // signed step of |y|'s mantissa
        // steps left before the binade boundary = floor(room / |D|).  Both are < 2^24, so the
        // correctly rounded float quotient is off by at most one: far cheaper than integer division.
        // Going down, stay strictly above the binade's bottom: an exact sum just below 2^e would
        // round on the finer grid of the binade underneath.
        int32_t m = n;
        if (D != 0) {
          const int32_t room = D > 0 ? (0x00FFFFFF - X) : (X - 0x00800001);
          const int32_t Da = D > 0 ? D : -D;
          m = room < 0 ? 0 : (int32_t)__fdiv_rn((float)room, (float)Da);
          if (m * Da > room) m--;
          if ((m + 1) * Da <= room) m++;
        }
        if (m > n) m = n;
        if (m > 0) {
          const int32_t X2 = X + m * D;
          x = __uint_as_float((by & 0xFF800000u) | ((uint32_t)X2 & 0x007FFFFFu));
          n -= m;
          continue;
        }
      }
    }
    x = y;
  }
  return x;
}
