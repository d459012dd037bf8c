/*
 * pioran_oracle_grad_ld.c — the gradient restatement (pioran_oracle_grad.c) compiled in x87 80-bit arithmetic.
 * TEST INFRASTRUCTURE ONLY.  Exports orc_approx_logl_grad_batch_ld: conditioning triage for gradient comparisons.
 */
#define ORC_GRAD_LONG_DOUBLE 1
#include "pioran_oracle_grad.c"
