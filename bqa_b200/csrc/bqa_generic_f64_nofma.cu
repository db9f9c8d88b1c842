// complex128 instantiation of the shape-generic kernels, compiled with -fmad=false (see build.py).
// complex128 is the validation precision (the reference default, src/bqa/utils.py:9-20).  On instances with
// exactly degenerate spectra (+-1 couplings: heavy-hex, MaxCut) the truncation picks vectors inside a
// degenerate singular subspace; which ones depends on whether exact zeros of the extended messages survive
// the arithmetic.  Separately rounded multiply and add (what numpy/LAPACK on the host do) keeps them, fused
// multiply-add does not -- so this translation unit must not contract.
#include "bqa_generic.cuh"
#include "bqa_multiclass.cuh"
namespace bqa {
BQA_INSTANTIATE(double)
BQA_INSTANTIATE_MULTICLASS(double)
}
