// Host-side TMA descriptor encoding.  cuTensorMapEncodeTiled lives in libcuda; it is resolved at run time through the
// runtime's driver entry-point query so the library links against cudart only.
#include "sky_common.cuh"

namespace sky {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int encode_nhwc_tensor_map(CUtensorMap *out, const float *base, int B, int h, int w, int C, int box_c, int box_w, int box_h)
{
    EncodeTiledFn enc = get_encoder();
    SKY_REQUIRE(enc != nullptr, SKY_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = { (cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B };
    cuuint64_t strides[3] = { (cuuint64_t)C * 4, (cuuint64_t)w * C * 4, (cuuint64_t)h * w * C * 4 };
    cuuint32_t box[4] = { (cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1 };
    cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SKY_REQUIRE(r == CUDA_SUCCESS, SKY_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (B=%d h=%d w=%d C=%d box=%dx%dx%d)", (int)r,
                B, h, w, C, box_c, box_w, box_h);
    return SKY_OK;
}

// dy [B, OH, OW, F] as the B operand of the strip weight gradient: dimensions ordered (f, panorama, column, row) so that a box of
// (32 filters, 8 panoramas, 8 columns, 1 row) lands in shared memory as rows = column * 8 + panorama of 128 bytes, in the 32-byte-atom
// 128-byte swizzle the MN-major TF32 operand descriptor expects (SWIZZLE_128B_BASE32B); out-of-range panoramas / columns / filters are zero.
int encode_dy_wgrad_tensor_map(CUtensorMap *out, const float *base, int B, int OH, int OW, int F, int box_cols)
{
    EncodeTiledFn enc = get_encoder();
    SKY_REQUIRE(enc != nullptr, SKY_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = { (cuuint64_t)F, (cuuint64_t)B, (cuuint64_t)OW, (cuuint64_t)OH };
    cuuint64_t strides[3] = { (cuuint64_t)OH * OW * F * 4, (cuuint64_t)F * 4, (cuuint64_t)OW * F * 4 };
    cuuint32_t box[4] = { 32, 8, (cuuint32_t)box_cols, 1 };
    cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SKY_REQUIRE(r == CUDA_SUCCESS, SKY_ERR_CUDA, "cuTensorMapEncodeTiled (dy, weight gradient) failed with CUresult %d (B=%d OH=%d OW=%d F=%d)", (int)r,
                B, OH, OW, F);
    return SKY_OK;
}

int encode_2d_tensor_map(CUtensorMap *out, const float *base, long rows, long cols, int box_cols, int box_rows)
{
    return encode_2d_tensor_map_sw(out, base, rows, cols, box_cols, box_rows, 0);
}

// swizzle128 != 0: box_cols must be 32 (128-byte rows); 16-byte chunk c of tile row r lands at chunk c ^ (r & 7)
int encode_2d_tensor_map_sw(CUtensorMap *out, const float *base, long rows, long cols, int box_cols, int box_rows, int swizzle128)
{
    EncodeTiledFn enc = get_encoder();
    SKY_REQUIRE(enc != nullptr, SKY_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = { (cuuint64_t)cols, (cuuint64_t)rows };
    cuuint64_t strides[1] = { (cuuint64_t)cols * 4 };
    cuuint32_t box[2] = { (cuuint32_t)box_cols, (cuuint32_t)box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SKY_REQUIRE(r == CUDA_SUCCESS, SKY_ERR_CUDA, "cuTensorMapEncodeTiled (2-D) failed with CUresult %d (rows=%ld cols=%ld box=%dx%d)", (int)r,
                rows, cols, box_rows, box_cols);
    return SKY_OK;
}

}  // namespace sky
