// hs_stream.cu -- host-side dispatch of the temporally blocked streaming kernel (hs_stream.cuh).  The kernel is
// instantiated per block depth T in hs_stream_inst.cu (one object per T).
#include <algorithm>

#include "hs_stream.cuh"

namespace hs {

#define HS_DECL(k)                                                                                                     \
    cudaError_t stream_prep_T##k();                                                                                    \
    cudaError_t stream_launch_T##k(int, bool, const CUtensorMap&, const CUtensorMap&, const StreamArgs&, cudaStream_t); \
    int stream_occ_T##k(int);
HS_DECL(1) HS_DECL(2) HS_DECL(3) HS_DECL(4) HS_DECL(5) HS_DECL(6) HS_DECL(7) HS_DECL(8)
#undef HS_DECL

cudaError_t stream_launch_track(int, const CUtensorMap&, const CUtensorMap&, const StreamArgs&, cudaStream_t);   // hs_stream_inst.cu, T = kTrackT

template <int T> static StreamGeom geom_of() {
    using C = typename DefaultCfg<T>::type;
    return StreamGeom{C::HL, C::VALIDW, C::SMEM_WARP, kMaxT, C::RG};
}
StreamGeom stream_geometry(int T) {
    switch (T) {
        case 1: return geom_of<1>(); case 2: return geom_of<2>(); case 3: return geom_of<3>(); case 4: return geom_of<4>();
        case 5: return geom_of<5>(); case 6: return geom_of<6>(); case 7: return geom_of<7>(); default: return geom_of<8>();
    }
}

cudaError_t stream_prepare(int) {
    cudaError_t e;
    if ((e = stream_prep_T1()) != cudaSuccess) return e;
    if ((e = stream_prep_T2()) != cudaSuccess) return e;
    if ((e = stream_prep_T3()) != cudaSuccess) return e;
    if ((e = stream_prep_T4()) != cudaSuccess) return e;
    if ((e = stream_prep_T5()) != cudaSuccess) return e;
    if ((e = stream_prep_T6()) != cudaSuccess) return e;
    if ((e = stream_prep_T7()) != cudaSuccess) return e;
    return stream_prep_T8();
}

int stream_warps_per_sm(int T, int stencil) {
    switch (T) {
        case 1: return stream_occ_T1(stencil); case 2: return stream_occ_T2(stencil);
        case 3: return stream_occ_T3(stencil); case 4: return stream_occ_T4(stencil);
        case 5: return stream_occ_T5(stencil); case 6: return stream_occ_T6(stencil);
        case 7: return stream_occ_T7(stencil); default: return stream_occ_T8(stencil);
    }
}

cudaError_t launch_jacobi_stream_track(int stencil, const CUtensorMap& tuv, const CUtensorMap& tc, StreamArgs A, int pairs,
                                       cudaStream_t s) {
    const StreamGeom G = stream_geometry(kTrackT);
    const int rows = A.out_hi - A.out_lo;
    if (rows <= 0 || pairs <= 0) return cudaSuccess;
    if (A.done_counter != nullptr || A.stop == nullptr || A.emax == nullptr || A.emax_next == nullptr || A.trk_t < 1 || A.trk_t > kTrackT)
        return cudaErrorInvalidValue;
    A.nsx = (A.W + G.valid_w - 1) / G.valid_w;
    A.ncy = (rows + A.chunk_rows - 1) / A.chunk_rows;
    A.total_units = (long long)A.nsx * A.ncy * pairs;
    A.seam_first = 0;
    A.signal_units = A.total_units;
    return stream_launch_track(stencil, tuv, tc, A, s);
}

bool stream_seam_first(int T, const StreamArgs& A) {
    const int rows = A.out_hi - A.out_lo;
    if (A.done_counter == nullptr || rows <= 0 || A.chunk_rows <= 0) return false;
    const int ncy = (rows + A.chunk_rows - 1) / A.chunk_rows;
    if (ncy < 2 || A.chunk_rows < T) return false;
    // Rows [0, up_lo) and [dn_hi, H) are ghost rows the neighbours store into; [up_lo, up_hi) and [dn_lo, dn_hi)
    // are the rows this strip pushes.  Early signalling is safe when only the first and the last chunk touch them.
    const int first_end = A.out_lo + A.chunk_rows, last_start = A.out_lo + (ncy - 1) * A.chunk_rows;
    const bool top_ok = A.peer_up == nullptr || first_end >= std::max(A.up_hi, A.up_lo + T);
    const bool bot_ok = A.peer_dn == nullptr || last_start <= std::min(A.dn_lo, A.dn_hi - T);
    return top_ok && bot_ok;
}

cudaError_t launch_jacobi_stream(int T, int stencil, const CUtensorMap& tuv, const CUtensorMap& tc, StreamArgs A, int pairs,
                                 cudaStream_t s) {
    const StreamGeom G = stream_geometry(T);
    const int rows = A.out_hi - A.out_lo;
    if (rows <= 0 || pairs <= 0) return cudaSuccess;
    A.nsx = (A.W + G.valid_w - 1) / G.valid_w;
    A.ncy = (rows + A.chunk_rows - 1) / A.chunk_rows;
    A.total_units = (long long)A.nsx * A.ncy * pairs;
    A.seam_first = 0;
    A.signal_units = A.total_units;
    if (stream_seam_first(T, A)) {
        A.seam_first = 1;
        A.signal_units = (long long)A.nsx * 2 * pairs;
    }
    const bool peer = A.done_counter != nullptr;    // strip connected to its neighbours (hsflow_strip_connect)
    switch (T) {
        case 1: return stream_launch_T1(stencil, peer, tuv, tc, A, s);
        case 2: return stream_launch_T2(stencil, peer, tuv, tc, A, s);
        case 3: return stream_launch_T3(stencil, peer, tuv, tc, A, s);
        case 4: return stream_launch_T4(stencil, peer, tuv, tc, A, s);
        case 5: return stream_launch_T5(stencil, peer, tuv, tc, A, s);
        case 6: return stream_launch_T6(stencil, peer, tuv, tc, A, s);
        case 7: return stream_launch_T7(stencil, peer, tuv, tc, A, s);
        case 8: return stream_launch_T8(stencil, peer, tuv, tc, A, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace hs
