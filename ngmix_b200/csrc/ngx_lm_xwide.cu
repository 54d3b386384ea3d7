// The multi-band LM kernels with a 20-parameter solver workspace: 8 to 12
// bands (ngx_lm_fit_multi_xwide; 5 to 7 shared parameters + a flux per band).
// Same source as ngx_lm.cu / ngx_lm_wide.cu: the wide entry points and their
// record type are compiled under the xwide names.
#include "../../include/ngx.h"
#define LM_MAXP 20
#define NGX_LM_WIDE_TU 1
#define NGX_LM_XWIDE_TU 1
#define ngx_lm_result_wide ngx_lm_result_xwide
#define ngx_lm_fit_multi_wide ngx_lm_fit_multi_xwide
#define ngx_lm_fit_multi_wide_host ngx_lm_fit_multi_xwide_host
#include "ngx_lm.cu"
