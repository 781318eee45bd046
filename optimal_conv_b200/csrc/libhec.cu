// libhec.cu -- single translation unit of libhec.so (kernels are defined in headers).
#include "hec.cu"
#include "hec_conv.cu"
#include "hec_poly.cu"
#include "hec_lt.cu"
#include "hec_encode.cu"
#include "hec_layer.cu"
