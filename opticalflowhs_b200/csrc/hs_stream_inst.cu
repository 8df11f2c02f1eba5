// hs_stream_inst.cu -- instantiates k_jacobi_stream for ONE temporal-block depth (-DHS_STREAM_T=k, k = 1..8), both
// stencils, with and without the peer transport.  Compiled once per depth so that the build runs in parallel.
#include "hs_stream.cuh"

#ifndef HS_STREAM_T
#error "compile with -DHS_STREAM_T=<1..8>"
#endif

namespace hs {

#define HS_CAT2(a, b) a##b
#define HS_CAT(a, b) HS_CAT2(a, b)
#define HS_FN(name) HS_CAT(name, HS_STREAM_T)

constexpr int kT = HS_STREAM_T;

template <int ST, bool PEER> static cudaError_t prep_one() {
    return cudaFuncSetAttribute(k_jacobi_stream<kT, ST, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
cudaError_t HS_FN(stream_prep_T)() {
    cudaError_t e;
    if ((e = prep_one<ST_CL8, false>()) != cudaSuccess) return e;
    if ((e = prep_one<ST_CL8, true>()) != cudaSuccess) return e;
    if ((e = prep_one<ST_CV4, false>()) != cudaSuccess) return e;
    return prep_one<ST_CV4, true>();
}

template <int ST, bool PEER>
static cudaError_t launch_one(const CUtensorMap& tuv, const CUtensorMap& tc, const StreamArgs& A, int wpc, cudaStream_t s) {
    using C = typename DefaultCfg<kT>::type;
    const size_t smem = (size_t)wpc * C::SMEM_WARP;
    const long long ctas = (A.total_units + wpc - 1) / wpc;
    if (ctas <= 0) return cudaSuccess;
    if (ctas > 0x7fffffffLL || smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    k_jacobi_stream<kT, ST, PEER><<<(unsigned)ctas, wpc * 32, smem, s>>>(tuv, tc, A);
    return cudaGetLastError();
}
cudaError_t HS_FN(stream_launch_T)(int st, bool peer, const CUtensorMap& tuv, const CUtensorMap& tc, const StreamArgs& A, int wpc,
                                   cudaStream_t s) {
    if (st == ST_CL8) return peer ? launch_one<ST_CL8, true>(tuv, tc, A, wpc, s) : launch_one<ST_CL8, false>(tuv, tc, A, wpc, s);
    return peer ? launch_one<ST_CV4, true>(tuv, tc, A, wpc, s) : launch_one<ST_CV4, false>(tuv, tc, A, wpc, s);
}

int HS_FN(stream_occ_T)(int st, int wpc) {
    using C = typename DefaultCfg<kT>::type;
    int n = 0;
    const size_t smem = (size_t)wpc * C::SMEM_WARP;
    cudaError_t e = st == ST_CL8 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_jacobi_stream<kT, ST_CL8, false>, wpc * 32, smem)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_jacobi_stream<kT, ST_CV4, false>, wpc * 32, smem);
    if (e != cudaSuccess) { cudaGetLastError(); return 8; }
    return n * wpc;
}

}  // namespace hs
