// The LM kernels with the 14-parameter solver workspace: coelliptical fits
// (ngx_lm_fit_coellip, 4 + 2 * ngauss parameters) and multi-band fits with 4
// to 7 bands (ngx_lm_fit_multi_wide).  Same source as ngx_lm.cu; a separate
// translation unit so that the 8-parameter kernels (the headline path) keep
// their smaller shared-memory state.
#define LM_MAXP 14
#define NGX_LM_WIDE_TU 1
#include "ngx_lm.cu"
