// fpb_curfit_d0.cu -- the fit pipeline (fpb_curfit.cu) instantiated for a run-time number (4..10) of
// coordinate curves per fit (parcur/fppara); a separate translation unit so that the
// K x weights x IDIM template grid compiles in parallel.
#define FPB_FIT_IDIM 0
#include "fpb_curfit.cu"
