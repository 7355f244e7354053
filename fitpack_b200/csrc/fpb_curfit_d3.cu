// fpb_curfit_d3.cu -- the fit pipeline (fpb_curfit.cu) instantiated for idim = 3
// coordinate curves per fit (parcur/fppara); a separate translation unit so that the
// K x weights x IDIM template grid compiles in parallel.
#define FPB_FIT_IDIM 3
#include "fpb_curfit.cu"
