/* TEST INFRASTRUCTURE.  Host build of the product's glibc_math.h next to the running libm, so that the
   restatement can be compared with libm bit for bit (tests/test_libm_exact.py).
   gcc -O2 -ffp-contract=off -mfma -shared -fPIC libm_check.c -lm      (-mfma: fma() must be one instruction,
   otherwise gcc calls libm's fma(), which is also exact — either way the result is the fused one). */
#include <math.h>
#include <stddef.h>
#include "../jitcsde_b200/csrc/glibc_math.h"

void restated_log_v(const double *x, double *out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = jsde_log(x[i]); }
void libm_log_v(const double *x, double *out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = log(x[i]); }
void restated_pow_v(const double *x, double y, double *out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = jsde_pow(x[i], y); }
void libm_pow_v(const double *x, double y, double *out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = pow(x[i], y); }
void restated_exp_v(const double *x, double *out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = jsde_exp(x[i]); }
void libm_exp_v(const double *x, double *out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = exp(x[i]); }
