"""The per-sample simplex projection divides by the size of the support (1..8).  The kernels
replace that IEEE division by q0 = a * RN(1/n), q = fma(fma(-n, q0, a), RN(1/n), q0)
(csrc/simplex.cuh, div_small_int), which must return the SAME double as a / n.  The same C
expression is compared here with the division on random operands (all exponents the solver
can meet, special mantissas included)."""

import os
import shutil
import subprocess

import pytest

SOURCE = r'''
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static uint64_t s = 88172645463325252ull;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(void)
{
    long bad = 0, tot = 0;
    for (int n = 1; n <= 8; ++n) {
        const double rn = 1.0 / (double)n;
        for (long i = 0; i < 4000000; ++i) {
            const uint64_t r = rnd();
            double a;
            if ((i & 3) == 0) {                    /* random mantissa, exponent in [-60, 60) */
                const uint64_t bits = (r & 0x800fffffffffffffull) |
                                      ((uint64_t)(1023 - 60 + (r >> 52) % 120) << 52);
                memcpy(&a, &bits, 8);
            } else if ((i & 3) == 1) {             /* few mantissa bits set */
                const uint64_t bits = (r & 0x800f00000000000full) | ((uint64_t)(1023 + (r >> 60)) << 52);
                memcpy(&a, &bits, 8);
            } else if ((i & 3) == 2) {             /* all-ones tails */
                const uint64_t bits = (r | 0x0000000000ffffffull) & 0x800fffffffffffffull;
                const uint64_t full = bits | ((uint64_t)(1023 - 3 + (r >> 61)) << 52);
                memcpy(&a, &full, 8);
            } else {
                a = ((double)(int64_t)r) * 0x1p-63 * 3.0 - 1.0;     /* (sum - 1) of weights */
            }
            const double q0 = a * rn;
            const double e = fma(-(double)n, q0, a);
            const double q = fma(e, rn, q0);
            if (q != a / (double)n) {
                if (bad < 5) printf("n=%d a=%a q=%a true=%a\n", n, a, q, a / (double)n);
                ++bad;
            }
            ++tot;
        }
    }
    printf("mismatches %ld of %ld\n", bad, tot);
    return bad != 0;
}
'''


@pytest.mark.skipif(shutil.which('gcc') is None, reason='needs gcc')
def test_small_integer_division_sequence_is_exact(tmp_path):
    src = tmp_path / 'divtest.c'
    src.write_text(SOURCE)
    exe = tmp_path / 'divtest'
    # -ffp-contract=off: the expression must be evaluated exactly as written
    subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-o', str(exe), str(src), '-lm'], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert 'mismatches 0 of 32000000' in out.stdout
