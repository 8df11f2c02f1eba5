// hs_stream_inst.cu -- instantiates k_jacobi_stream for ONE temporal-block depth (-DHS_STREAM_T=k, k = 1..8), both
// stencils, with and without the peer transport.  Compiled once per depth so that the build runs in parallel.
#include <cstdlib>

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

// Every instantiation is launched with programmatic stream serialisation (the kernel calls griddepcontrol.wait before
// it reads anything a previous launch wrote); HSFLOW_NO_PDL=1 switches the attribute off for A/B measurements.
static bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("HSFLOW_NO_PDL"); return !(e && atoi(e)); }();
    return on;
}
template <typename K>
static cudaError_t launch_pdl(K kernel, unsigned ctas, size_t smem, cudaStream_t s, const CUtensorMap& tuv, const CUtensorMap& tc,
                              const StreamArgs& A) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, tuv, tc, A);
}

template <int ST, bool PEER>
static cudaError_t launch_one(const CUtensorMap& tuv, const CUtensorMap& tc, const StreamArgs& A, cudaStream_t s) {
    using C = typename DefaultCfg<kT>::type;
    const long long ctas = A.total_units;              // one autonomous warp per CTA
    if (ctas <= 0) return cudaSuccess;
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    return launch_pdl(k_jacobi_stream<kT, ST, PEER>, (unsigned)ctas, C::SMEM_WARP, s, tuv, tc, A);
}
cudaError_t HS_FN(stream_launch_T)(int st, bool peer, const CUtensorMap& tuv, const CUtensorMap& tc, const StreamArgs& A,
                                   cudaStream_t s) {
    if (st == ST_CL8) return peer ? launch_one<ST_CL8, true>(tuv, tc, A, s) : launch_one<ST_CL8, false>(tuv, tc, A, s);
    return peer ? launch_one<ST_CV4, true>(tuv, tc, A, s) : launch_one<ST_CV4, false>(tuv, tc, A, s);
}

#if HS_STREAM_T == 4
// The EPS-criterion (TRACK) instantiation exists for this depth only (kTrackT).
static_assert(kTrackT == 4, "instantiate k_jacobi_stream<kTrackT, .., TRACK = true> in the matching object");
template <int ST> static cudaError_t launch_track_one(const CUtensorMap& tuv, const CUtensorMap& tc, const StreamArgs& A, cudaStream_t s) {
    using C = typename DefaultCfg<kT>::type;
    static bool prepared = false;                  // one attribute call per instantiation and process
    if (!prepared) {
        cudaError_t e = cudaFuncSetAttribute(k_jacobi_stream<kT, ST, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        prepared = true;
    }
    if (A.total_units <= 0) return cudaSuccess;
    if (A.total_units > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    return launch_pdl(k_jacobi_stream<kT, ST, false, true>, (unsigned)A.total_units, C::SMEM_WARP, s, tuv, tc, A);
}
cudaError_t stream_launch_track(int st, const CUtensorMap& tuv, const CUtensorMap& tc, const StreamArgs& A, cudaStream_t s) {
    return st == ST_CL8 ? launch_track_one<ST_CL8>(tuv, tc, A, s) : launch_track_one<ST_CV4>(tuv, tc, A, s);
}
#endif

int HS_FN(stream_occ_T)(int st) {
    using C = typename DefaultCfg<kT>::type;
    int n = 0;
    cudaError_t e = st == ST_CL8 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_jacobi_stream<kT, ST_CL8, false>, 32, C::SMEM_WARP)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_jacobi_stream<kT, ST_CV4, false>, 32, C::SMEM_WARP);
    if (e != cudaSuccess) { cudaGetLastError(); return 8; }
    return n;
}

}  // namespace hs
