// fpb_percur_solo.cu -- percur for a FEW curves (fp_percur_c, small batches): the per-curve code of fpb_percur.cu
// compiled a second time with its workspace in shared memory, one curve per CTA, the whole of fpperi
// (core:9668-10192) in one launch -- see the FPB_PER_SOLO notes in fpb_percur.cu.
#define FPB_PER_SOLO 1
#include "fpb_percur.cu"
