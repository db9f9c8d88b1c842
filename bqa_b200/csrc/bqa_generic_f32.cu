// complex64 instantiation of the shape-generic kernels (FMA contraction on).
#include "bqa_generic.cuh"
#include "bqa_multiclass.cuh"
namespace bqa {
BQA_INSTANTIATE(float)
BQA_INSTANTIATE_MULTICLASS(float)
}
