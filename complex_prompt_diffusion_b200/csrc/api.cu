// Error plumbing, ABI version and the host-side TMA tensor-map encoder.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/cpd_b200.h"
#include "common.cuh"

static thread_local char g_err[1024] = "";

void cpd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* cpd_last_error(void) { return g_err; }

bool cpd_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CPD_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
extern "C" int cpd_abi_version(void) { return 1; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int cpd_make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box) {
  return cpd_make_tmap16(map, base, rank, dims, strides_bytes, box, 128);
}

int cpd_make_tmap16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    cpd_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return CPD_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if (((uintptr_t)base & 15) != 0) {
    cpd_set_error("tensor map base %p is not 16-byte aligned", base);
    return CPD_ERR_INVALID;
  }
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) {
      cpd_set_error("tensor map stride[%d] = %llu bytes is not a multiple of 16", i, (unsigned long long)gstr[i]);
      return CPD_ERR_INVALID;
    }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, gbox,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                        : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                               : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE)),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cpd_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,.. box %u,%u,..)", (int)r, rank,
                  (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0), gbox[0],
                  rank > 1 ? gbox[1] : 0);
    return CPD_ERR_CUDA;
  }
  return CPD_OK;
}
