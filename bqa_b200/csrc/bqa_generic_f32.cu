// complex64 instantiation of the shape-generic kernels (FMA contraction on).
#include "bqa_generic.cuh"
namespace bqa {
BQA_INSTANTIATE(float)
}
