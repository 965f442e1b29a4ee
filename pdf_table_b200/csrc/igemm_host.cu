// Host side of conv_igemm_tcgen05: TMA tensor-map construction, tile planning, launch.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "engine.h"
#include "igemm.cuh"
#include "win_conv.cuh"
#include "mlp_fused.cuh"
#include "dcn_fused.cuh"

namespace dv {

// cuTensorMapEncodeTiled is a driver entry point; resolve it through the runtime so the library
// does not link against libcuda (which is absent on the CPU-only build box).
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

static int encode_map(Engine* e, CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, int row_bytes,
                      const char* what) {
    auto fn = get_encode_fn();
    if (!fn) return set_err(e, DV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = row_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                    gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        return set_err(e, DV_ERR_CUDA,
                       "cuTensorMapEncodeTiled(%s) failed rc=%d dims={%llu,%llu,%llu,%llu,%llu} "
                       "strides={%llu,%llu,%llu,%llu} box={%u,%u,%u,%u,%u} rank=%d base=%p",
                       what, (int)r, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                       (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
                       (unsigned long long)(rank > 4 ? gdim[4] : 0), (unsigned long long)gstr[0],
                       (unsigned long long)(rank > 2 ? gstr[1] : 0), (unsigned long long)(rank > 3 ? gstr[2] : 0),
                       (unsigned long long)(rank > 4 ? gstr[3] : 0), bdim[0], rank > 1 ? bdim[1] : 0,
                       rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0, rank, base);
    }
    return 0;
}

static void choose_patch(int Ho, int Wo, int* TH, int* TW) {
    long long best = -1;
    int bth = 8, btw = 16;
    for (int th = 1; th <= 128; th *= 2) {
        const int tw = 128 / th;
        const long long tiles = static_cast<long long>((Ho + th - 1) / th) * ((Wo + tw - 1) / tw);
        // prefer fewer tiles; tie-break towards squarer patches (less halo re-read through L2)
        const long long cost = tiles * 1000 + (th > tw ? th / tw : tw / th);
        if (best < 0 || cost < best) {
            best = cost;
            bth = th;
            btw = tw;
        }
    }
    *TH = bth;
    *TW = btw;
}

static int finish_plan(Engine* e, ConvPlan* plan, const ConvSpec& cs, const EpiSpec& es, int num_kb,
                       const std::vector<int4>& deltas, const char* name) {
    IGemmParams& p = plan->prm;
    p.num_kb = num_kb;
    p.BK = cs.BK;
    p.Cout = cs.Cout;
    int block_n;
    if (cs.Cout <= 256) {
        block_n = (cs.Cout + 31) / 32 * 32;
    } else {
        // 256-wide n-tiles unless 128-wide ones waste fewer padded columns (360 -> 3 x 128; 480 -> 2 x 256: the A tile is re-read
        // per n-tile and a 128-wide MMA leaves the issuing thread as little time per k-block as it needs itself)
        const int waste256 = (cs.Cout + 255) / 256 * 256 - cs.Cout, waste128 = (cs.Cout + 127) / 128 * 128 - cs.Cout;
        block_n = (waste256 <= waste128 || cs.Cout > 1024) ? 256 : 128;
    }
    p.BLOCK_N = block_n;
    p.n_tiles = (cs.Cout + block_n - 1) / block_n;
    static const int acc_env = getenv("DV_ACC_STAGES") ? atoi(getenv("DV_ACC_STAGES")) : 0;
    p.acc_stages = block_n <= 64 ? 8 : block_n <= 128 ? 4 : 2;
    if (acc_env == 2 || acc_env == 4 || acc_env == 8) p.acc_stages = std::min(p.acc_stages, acc_env);
    const int row_bytes = 2 * cs.BK;
    // TMA-store epilogue: plain fp16 NHWC outputs (any A mode).  Not for fp32 / replicated / pixel-shuffled / split stores, the
    // arg-max epilogue or device-side row counts (the store would also write the rows past the count).
    static const bool tma_store_env = !(getenv("DV_TMA_STORE") && atoi(getenv("DV_TMA_STORE")) == 0);
    bool tma_store = tma_store_env && !es.out_f32 && es.out_mode == OUT_NHWC && es.split_off == 0 && es.arg_out == nullptr &&
                     es.m_dyn == nullptr && cs.Cout >= 32 && es.out != nullptr;
    size_t stg_bytes = tma_store ? 8 * 2 * 2048 : 0;
    size_t ring_budget = 206 * 1024 - stg_bytes;
    size_t ring_bytes = 0;
    if (p.mode == A_HALO) {
        // shared memory: halo-patch ring (18 x 16 pixels x BK channels each) + a ring of per-tap weight tiles
        const size_t halo_bytes = static_cast<size_t>(18) * 16 * row_bytes, b_bytes = static_cast<size_t>(block_n) * row_bytes;
        const size_t filter_bytes = 9 * static_cast<size_t>(p.ncb) * b_bytes;
        if (tma_store && p.n_tiles == 1 && filter_bytes <= 96 * 1024 && (ring_budget - filter_bytes) / halo_bytes < 3 &&
            (206 * 1024 - filter_bytes) / halo_bytes >= 3) {
            // a resident filter with a three-deep patch ring beats the store staging (measured on the 64 -> 64 layers)
            tma_store = false;
            stg_bytes = 0;
            ring_budget = 206 * 1024;
        }
        // small filters (one n-tile, <= 96 KB) stay resident for the whole kernel and the ring budget buys a deeper patch ring
        static const int min_hs = getenv("DV_HALO_MINHS") ? atoi(getenv("DV_HALO_MINHS")) : 3;
        if (p.n_tiles == 1 && filter_bytes <= 96 * 1024 && static_cast<int>((ring_budget - filter_bytes) / halo_bytes) >= min_hs) {
            int hs = static_cast<int>((ring_budget - filter_bytes) / halo_bytes);
            if (hs > kMaxHaloStages) hs = kMaxHaloStages;
            p.b_resident = 1;
            p.halo_stages = hs;
            p.num_stages = 1;
            ring_bytes = hs * halo_bytes + filter_bytes;
        } else {
            int hs = p.ncb > 1 ? 3 : 2;
            int stages = static_cast<int>((ring_budget - hs * halo_bytes) / b_bytes);
            if (stages < 3 && hs > 2) {
                hs = 2;
                stages = static_cast<int>((ring_budget - hs * halo_bytes) / b_bytes);
            }
            if (stages > kMaxStages) stages = kMaxStages;
            if (stages < 2) return set_err(e, DV_ERR_UNSUPPORTED, "%s: halo stage too large", name);
            p.halo_stages = hs;
            p.num_stages = stages;
            ring_bytes = hs * halo_bytes + stages * b_bytes;
        }
    } else {
        const size_t stage_bytes = static_cast<size_t>(128 + block_n) * row_bytes;
        int stages = static_cast<int>(ring_budget / stage_bytes);
        if (stages > kMaxStages) stages = kMaxStages;
        if (stages < 2) return set_err(e, DV_ERR_UNSUPPORTED, "%s: stage too large", name);
        p.num_stages = stages;
        ring_bytes = stages * stage_bytes;
    }
    p.tma_store = tma_store ? 1 : 0;
    p.stg_off = static_cast<int>((ring_bytes + 1023) / 1024 * 1024);
    plan->smem = p.stg_off + stg_bytes + 1024;
    if (tma_store) {
        const uint64_t ld_bytes = static_cast<uint64_t>(es.out_ld) * 2;
        const __half* obase = reinterpret_cast<const __half*>(es.out) + es.out_coff;
        if (p.mode == A_FLAT) {
            uint64_t dims[2] = {static_cast<uint64_t>(cs.Cout), static_cast<uint64_t>(p.M)}, str[1] = {ld_bytes};
            uint32_t box[2] = {32, 32};
            DV_TRY(encode_map(e, &p.tmD, obase, 2, dims, str, box, 64, name));
        } else {
            p.st_bw = p.TW >= 32 ? 32 : p.TW;
            if (32 % p.st_bw) return set_err(e, DV_ERR_UNSUPPORTED, "%s: tile width %d does not divide a warp's 32 rows", name, p.TW);
            uint64_t dims[4] = {static_cast<uint64_t>(cs.Cout), static_cast<uint64_t>(p.Wo), static_cast<uint64_t>(p.Ho), static_cast<uint64_t>(p.Nimg)};
            uint64_t str[3] = {ld_bytes, ld_bytes * p.Wo, ld_bytes * p.Wo * p.Ho};
            uint32_t box[4] = {32, static_cast<uint32_t>(p.st_bw), static_cast<uint32_t>(32 / p.st_bw), 1};
            DV_TRY(encode_map(e, &p.tmD, obase, 4, dims, str, box, 64, name));
        }
    }
    p.bias = cs.bias;
    p.res = es.res;
    p.res_mode = es.res_mode;
    p.res_ld = es.res_ld;
    p.res_mod = es.res_mod;
    p.arg_out = es.arg_out;
    p.max_out = es.max_out;
    p.n_inner = es.arg_out != nullptr ? 1 : 0;
    plan->res_f32 = es.res_f32;
    if (es.arg_out != nullptr && !es.out_f32) return set_err(e, DV_ERR_ARG, "%s: arg-max epilogue needs out_f32", name);
    if (es.res_f32 && es.res_mode != RES_NONE && (es.res_ld % 4))
        return set_err(e, DV_ERR_UNSUPPORTED, "%s: fp32 residual needs res_ld %% 4 == 0", name);
    p.act = es.act;
    p.out_mode = es.out_mode;
    p.out = es.out;
    p.out_ld = es.out_ld;
    p.out_coff = es.out_coff;
    p.rep = es.rep;
    p.out_f32 = es.out_f32;
    p.m_dyn = es.m_dyn;
    p.split_off = es.split_off;
    p.post_affine = es.post_affine;
    p.post_scale = es.post_scale;
    p.post_bias = es.post_bias;
    p.res_split_off = es.res_lo;
    if (es.res_lo && (es.res_f32 || es.res_mode == RES_NONE || (es.res_lo % 8)))
        return set_err(e, DV_ERR_ARG, "%s: split residual needs an fp16 residual and res_lo %% 8 == 0", name);
    p.res_red = (es.res_f32 && es.out_f32 && es.res != nullptr && es.res == es.out && es.res_mode == RES_SAME && es.res_mod == 0 &&
                 es.res_ld == es.out_ld && es.out_coff == 0 && es.out_mode == OUT_NHWC && es.act == ACT_NONE && es.arg_out == nullptr &&
                 !(getenv("DV_RES_RED") && atoi(getenv("DV_RES_RED")) == 0))
                    ? 1 : 0;
    if (es.m_dyn && p.mode != A_FLAT) return set_err(e, DV_ERR_ARG, "%s: m_dyn needs a flat GEMM", name);
    if (es.split_off && (es.out_f32 || (es.split_off % 8)))
        return set_err(e, DV_ERR_ARG, "%s: split store needs fp16 output and split_off %% 8 == 0", name);
    if (num_kb > kMaxKB) return set_err(e, DV_ERR_UNSUPPORTED, "%s: %d k-blocks > %d", name, num_kb, kMaxKB);
    {
        const int al = es.out_f32 ? 4 : 8;
        if ((cs.Cout % al) || (es.out_ld % al) || (es.out_coff % al) || (es.res_mode != RES_NONE && (es.res_ld % 8)))
            return set_err(e, DV_ERR_UNSUPPORTED,
                           "%s: Cout/out_ld/out_coff must be multiples of %d and res_ld of 8 (pad channels)", name, al);
        if (es.out == nullptr && es.arg_out == nullptr) return set_err(e, DV_ERR_ARG, "%s: no output", name);
        if ((reinterpret_cast<uintptr_t>(es.out) & 15) || (reinterpret_cast<uintptr_t>(es.res) & 15))
            return set_err(e, DV_ERR_ARG, "%s: out/res pointers must be 16-byte aligned", name);
    }
    if (es.out_mode == OUT_SHUF2 && (((cs.Cout >> 2) % 32) != 0 || (cs.Cout & 3)))
        return set_err(e, DV_ERR_UNSUPPORTED, "%s: OUT_SHUF2 needs Cout/4 %% 32 == 0", name);
    // weight map: [Cout][Ktot] K-major
    const uint64_t ktot = static_cast<uint64_t>(num_kb) * cs.BK;
    {
        uint64_t dims[2] = {ktot, static_cast<uint64_t>(cs.Cout)};
        uint64_t str[1] = {ktot * 2};
        uint32_t box[2] = {static_cast<uint32_t>(cs.BK), static_cast<uint32_t>(block_n)};
        DV_TRY(encode_map(e, &p.tmB, cs.w, 2, dims, str, box, row_bytes, name));
    }
    void* dtab = nullptr;
    DV_TRY(e->dalloc(&dtab, deltas.size() * sizeof(int4)));
    DV_CUDA(e, cudaMemcpyAsync(dtab, deltas.data(), deltas.size() * sizeof(int4), cudaMemcpyHostToDevice,
                               e->stream));
    DV_CUDA(e, cudaStreamSynchronize(e->stream));
    p.kb_delta = reinterpret_cast<const int4*>(dtab);
    {
        // algorithmic HBM bytes: A once, W once, D once (fp16/fp32), residual once
        const double m_rows = p.mode == A_FLAT ? (double)p.M : (double)p.Nimg * p.Ho * p.Wo;
        double out_elems = m_rows * cs.Cout;
        if (es.out_mode == OUT_REPL) out_elems *= (double)es.rep * es.rep;
        plan->bytes += (double)ktot * cs.Cout * 2 + out_elems * (es.out_f32 ? 4 : 2);
        if (es.arg_out != nullptr) plan->bytes += m_rows * 8 - (es.out == nullptr ? out_elems * 4 : 0);
        if (es.split_off) plan->bytes += out_elems * 2;
        if (es.res_mode == RES_SAME) plan->bytes += (es.res_mod > 0 ? (double)es.res_mod : m_rows) * cs.Cout * (es.res_f32 ? 4 : (es.res_lo ? 4 : 2));
        if (es.res_mode == RES_UP2) plan->bytes += m_rows * cs.Cout * (es.res_lo ? 4 : 2) / 4;
    }
    const int total = p.n_inner ? p.m_tiles : p.m_tiles * p.n_tiles;
    plan->grid = total < e->num_sms ? total : e->num_sms;
    plan->name = name;
    return 0;
}

int plan_linear(Engine* e, const __half* A, int M, int K, const ConvSpec& cs, const EpiSpec& es,
                ConvPlan* plan, const char* name, int lda, int lo_off) {
    memset(&plan->prm, 0, sizeof(plan->prm));
    plan->bytes = 0;
    IGemmParams& p = plan->prm;
    if (K % 8) return set_err(e, DV_ERR_UNSUPPORTED, "%s: K %% 8 != 0", name);
    // split-fp16: A holds [hi | lo] (lo_off columns apart, K by default); the k-block list walks hi, hi, lo against
    // W = [W_hi | W_lo | W_hi]
    if (cs.split && lo_off == 0) lo_off = K;
    if (cs.split && (lo_off < K || (lo_off % 8))) return set_err(e, DV_ERR_ARG, "%s: bad lo offset %d", name, lo_off);
    const int a_cols = cs.split ? lo_off + K : K;
    if (lda == 0) lda = a_cols;
    if ((lda % 8) || lda < a_cols || (reinterpret_cast<uintptr_t>(A) & 15))
        return set_err(e, DV_ERR_UNSUPPORTED, "%s: A row stride %d / alignment", name, lda);
    p.mode = A_FLAT;
    p.M = M;
    p.Nimg = 1;
    p.Ho = 1;
    p.Wo = M;
    p.TH = 1;
    p.TW = 128;
    p.tiles_x = p.tiles_y = 1;
    p.m_tiles = (M + 127) / 128;
    const int kb_per = (K + cs.BK - 1) / cs.BK;
    if (cs.split && (K % cs.BK)) return set_err(e, DV_ERR_UNSUPPORTED, "%s: split GEMM needs BK | K", name);
    const int num_kb = cs.split ? 3 * kb_per : kb_per;
    if (cs.Cin_pad != num_kb * cs.BK)
        return set_err(e, DV_ERR_WEIGHTS, "%s: weight packing Cin_pad=%d != %d", name, cs.Cin_pad, num_kb * cs.BK);
    {
        uint64_t dims[5] = {static_cast<uint64_t>(a_cols), static_cast<uint64_t>(M), 1, 1, 1};
        const uint64_t rs = static_cast<uint64_t>(lda) * 2;
        uint64_t str[4] = {rs, rs * M, rs * M, rs * M};
        uint32_t box[5] = {static_cast<uint32_t>(cs.BK), 128, 1, 1, 1};
        DV_TRY(encode_map(e, &p.tmA, A, 5, dims, str, box, 2 * cs.BK, name));
    }
    std::vector<int4> deltas(num_kb);
    for (int kb = 0; kb < num_kb; ++kb) {
        int col = kb * cs.BK;
        if (cs.split) {
            const int part = kb / kb_per, k = (kb - part * kb_per) * cs.BK;
            col = part == 2 ? lo_off + k : k;
        }
        deltas[kb] = make_int4(col, 0, 0, 0);
    }
    plan->flops = 2.0 * M * K * cs.Cout * (cs.split ? 3 : 1);
    plan->bytes = 2.0 * M * a_cols;
    return finish_plan(e, plan, cs, es, num_kb, deltas, name);
}

int plan_conv(Engine* e, const Tensor& in, const ConvSpec& cs, const EpiSpec& es, int Ho, int Wo,
              ConvPlan* plan, const char* name) {
    if (cs.flat || (cs.KH == 1 && cs.KW == 1 && cs.stride == 1 && cs.pad == 0 && !cs.stem)) {
        if (cs.split != (in.lo > 0)) return set_err(e, DV_ERR_ARG, "%s: split weights need a split input (and vice versa)", name);
        int rc = plan_linear(e, in.p, in.N * in.H * in.W, in.C, cs, es, plan, name, in.ldc(), static_cast<int>(in.lo));
        if (rc == 0) {
            plan->prm.Nimg = in.N;
            plan->prm.Ho = in.H;
            plan->prm.Wo = in.W;
        }
        return rc;
    }
    memset(&plan->prm, 0, sizeof(plan->prm));
    plan->bytes = 0;
    IGemmParams& p = plan->prm;
    p.Nimg = in.N;
    p.Ho = Ho;
    p.Wo = Wo;
    choose_patch(Ho, Wo, &p.TH, &p.TW);
    p.tiles_y = (Ho + p.TH - 1) / p.TH;
    p.tiles_x = (Wo + p.TW - 1) / p.TW;
    p.m_tiles = in.N * p.tiles_x * p.tiles_y;
    const int row_bytes = 2 * cs.BK;
    std::vector<int4> deltas;
    int num_kb = 0;
    if (cs.stem) {
        // `in` is the zero-bordered image [N, Hp, Wp, cpp]: stride 2 -> cpp = 4 channels per pixel, stride 1 -> 8, so
        // that one filter-tap row (8 pixels: 7 taps + 1 zero-weight pad) is a 16-byte-strided window of 8*cpp fp16.
        const int st = cs.stride, cpp = st == 2 ? 4 : 8, win = 8 * cpp;
        if ((st != 1 && st != 2) || cs.BK != win || in.C != cpp || in.ld != 0 || cs.KH != 7)
            return set_err(e, DV_ERR_UNSUPPORTED, "%s: stem expects 7x7, stride 1 (C=8, BK=64) or 2 (C=4, BK=32)", name);
        p.mode = A_STEM;
        const uint64_t pitch = static_cast<uint64_t>(in.W) * cpp * 2;
        if (in.W < st * (Wo - 1) + 8 || in.H < st * (Ho - 1) + 7 || (pitch % 16))
            return set_err(e, DV_ERR_ARG, "%s: padded stem input too small", name);
        // fp32x: the lo copy of the padded image is a second batch of N images right behind the first
        if (cs.split != (in.lo > 0) || (in.lo && in.lo != static_cast<long long>(in.N) * in.H * in.W * cpp))
            return set_err(e, DV_ERR_ARG, "%s: split stem needs the lo images stacked behind the hi images", name);
        uint64_t dims[5] = {static_cast<uint64_t>(win), static_cast<uint64_t>(Wo), 7, static_cast<uint64_t>(Ho),
                            static_cast<uint64_t>(in.N) * (cs.split ? 2 : 1)};
        uint64_t str[4] = {16, pitch, static_cast<uint64_t>(st) * pitch, pitch * in.H};
        uint32_t box[5] = {static_cast<uint32_t>(win), static_cast<uint32_t>(p.TW), 1, static_cast<uint32_t>(p.TH), 1};
        DV_TRY(encode_map(e, &p.tmA, in.p, 5, dims, str, box, row_bytes, name));
        for (int r = 0; r < 7; ++r)
            for (int part = 0; part < (cs.split ? 3 : 1); ++part) deltas.push_back(make_int4(0, 0, r, part == 2 ? in.N : 0));
        num_kb = static_cast<int>(deltas.size());
        plan->flops = 2.0 * in.N * Ho * Wo * 147.0 * cs.Cout * (cs.split ? 3 : 1);
        plan->bytes = 2.0 * in.N * in.H * in.W * cpp * (cs.split ? 2 : 1);
    } else {
        if (in.C != cs.Cin) return set_err(e, DV_ERR_ARG, "%s: Cin mismatch %d vs %d", name, in.C, cs.Cin);
        if (in.C % 8) return set_err(e, DV_ERR_UNSUPPORTED, "%s: Cin %% 8 != 0", name);
        // cs.Cin_pad = packed channels per (tap, part): split weights hold 3 parts [W_hi | W_lo | W_hi] per tap
        const int cin_blocks = cs.Cin_pad / cs.BK;
        if (cin_blocks * cs.BK != cs.Cin_pad || cs.Cin_pad < cs.Cin)
            return set_err(e, DV_ERR_WEIGHTS, "%s: bad Cin_pad", name);
        const int parts = cs.split ? 3 : 1;
        const int lo = static_cast<int>(in.lo);
        if (cs.split != (in.lo > 0)) return set_err(e, DV_ERR_ARG, "%s: split weights need a split input (and vice versa)", name);
        if (cs.split && ((in.C % cs.BK) || (lo % 8) || lo < in.C || lo + in.C > in.ldc()))
            return set_err(e, DV_ERR_UNSUPPORTED, "%s: split conv needs BK | Cin and the lo half inside the pixel", name);
        const int c_extent = cs.split ? lo + in.C : in.C;  // channels the A map must reach from the slice start
        const uint64_t cb = static_cast<uint64_t>(in.ldc()) * 2;  // bytes between consecutive pixels
        if ((in.ldc() % 8) || (reinterpret_cast<uintptr_t>(in.p) & 15))
            return set_err(e, DV_ERR_UNSUPPORTED, "%s: input slice must be 16-byte aligned (ld %% 8, offset %% 8)", name);
        static const int halo_env = getenv("DV_HALO") ? atoi(getenv("DV_HALO")) : 2;  // streaming-filter halo mode (1) is slower than the patch mode (profiles/r1f, r4x)
        static const int halo_baseoff = getenv("DV_HALO_BASEOFF") ? atoi(getenv("DV_HALO_BASEOFF")) : 0;  // measured on B200: the swizzle is a function of the absolute shared-memory address, shifted starts need NO base offset
        // DV_HALO: 0 = never, 1 = every 3x3 stride-1 conv, 2 = only where the whole filter fits resident (<= 96 KB, one n-tile)
        const size_t filter_bytes = 9ull * cin_blocks * ((cs.Cout + 31) / 32 * 32) * 2 * cs.BK;
        const bool filter_fits = cs.Cout <= 256 && filter_bytes <= 96 * 1024 && (206 * 1024 - filter_bytes) / (288ull * 2 * cs.BK) >= 3;
        if (cs.stride == 1 && cs.KH == 3 && cs.KW == 3 && cs.pad == 1 && (halo_env == 1 || (halo_env == 2 && filter_fits)) && !cs.split) {
            // halo-patch mode: 16 x 8 output pixels per tile, one {BK, 16, 18} box per channel block
            p.mode = A_HALO;
            p.TH = 16;
            p.TW = 8;
            p.tiles_y = (Ho + p.TH - 1) / p.TH;
            p.tiles_x = (Wo + p.TW - 1) / p.TW;
            p.m_tiles = in.N * p.tiles_x * p.tiles_y;
            p.ncb = cin_blocks;
            p.desc_base_off = halo_baseoff;
            uint64_t dims[5] = {static_cast<uint64_t>(in.C), static_cast<uint64_t>(in.W),
                                static_cast<uint64_t>(in.H), static_cast<uint64_t>(in.N), 1};
            uint64_t str[4] = {cb, cb * in.W, cb * in.W * in.H, cb * in.W * in.H * in.N};
            uint32_t box[5] = {static_cast<uint32_t>(cs.BK), 16, 18, 1, 1};
            DV_TRY(encode_map(e, &p.tmA, in.p, 5, dims, str, box, row_bytes, name));
            for (int t = 0; t < 9 * cin_blocks; ++t) deltas.push_back(make_int4(0, 0, 0, 0));  // unused by the halo path
        } else if (cs.stride == 1) {
            p.mode = A_PATCH;
            uint64_t dims[5] = {static_cast<uint64_t>(c_extent), static_cast<uint64_t>(in.W),
                                static_cast<uint64_t>(in.H), static_cast<uint64_t>(in.N), 1};
            uint64_t str[4] = {cb, cb * in.W, cb * in.W * in.H, cb * in.W * in.H * in.N};
            uint32_t box[5] = {static_cast<uint32_t>(cs.BK), static_cast<uint32_t>(p.TW),
                               static_cast<uint32_t>(p.TH), 1, 1};
            DV_TRY(encode_map(e, &p.tmA, in.p, 5, dims, str, box, row_bytes, name));
            for (int r = 0; r < cs.KH; ++r)
                for (int s = 0; s < cs.KW; ++s)
                    for (int part = 0; part < parts; ++part)
                        for (int c = 0; c < cin_blocks; ++c)
                            deltas.push_back(make_int4((part == 2 ? lo : 0) + c * cs.BK, s - cs.pad, r - cs.pad, 0));
        } else if (cs.stride == 2) {
            if ((in.H & 1) || (in.W & 1) || (cs.Cin % cs.BK) || (in.ld != 0 && in.ld != c_extent))
                return set_err(e, DV_ERR_UNSUPPORTED, "%s: stride-2 needs even H,W, BK | Cin and a dense input", name);
            p.mode = A_PATCH_S2;
            const int pix_c = in.ldc();  // channels per pixel (2C for a split pair)
            uint64_t dims[5] = {static_cast<uint64_t>(2 * pix_c), static_cast<uint64_t>(in.W / 2), 2,
                                static_cast<uint64_t>(in.H / 2), static_cast<uint64_t>(in.N)};
            uint64_t str[4] = {2 * cb, cb * in.W, 2 * cb * in.W, cb * in.W * in.H};
            uint32_t box[5] = {static_cast<uint32_t>(cs.BK), static_cast<uint32_t>(p.TW), 1,
                               static_cast<uint32_t>(p.TH), 1};
            DV_TRY(encode_map(e, &p.tmA, in.p, 5, dims, str, box, row_bytes, name));
            for (int r = 0; r < cs.KH; ++r)
                for (int s = 0; s < cs.KW; ++s)
                    for (int part = 0; part < parts; ++part)
                        for (int c = 0; c < cin_blocks; ++c) {
                            const int xi = s - cs.pad, yi = r - cs.pad;
                            const int px = xi & 1, py = yi & 1;
                            deltas.push_back(make_int4(px * pix_c + (part == 2 ? lo : 0) + c * cs.BK, (xi - px) / 2, py, (yi - py) / 2));
                        }
        } else {
            return set_err(e, DV_ERR_UNSUPPORTED, "%s: stride %d", name, cs.stride);
        }
        num_kb = static_cast<int>(deltas.size());
        plan->flops = 2.0 * in.N * Ho * Wo * (double)cs.KH * cs.KW * cs.Cin * cs.Cout * parts;
        plan->bytes = 2.0 * in.N * in.H * in.W * in.C * (cs.split ? 2 : 1);
    }
    return finish_plan(e, plan, cs, es, num_kb, deltas, name);
}

typedef void (*IGemmKernel)(const IGemmParams);

// Instantiated epilogue variants: every activation x {fp16, fp32} output with an fp16 (or no) residual,
// plus the two fp32-stream specials (fp32 residual updated in place; classifier arg-max).
static IGemmKernel pick_kernel(int act, int out_f32, int res_f32, int argmax) {
    if (argmax) return (act == ACT_NONE && !res_f32) ? conv_igemm_tcgen05<ACT_NONE, true, false, true> : nullptr;
    if (res_f32) return (act == ACT_NONE && out_f32) ? conv_igemm_tcgen05<ACT_NONE, true, true, false> : nullptr;
    switch (act * 2 + (out_f32 ? 1 : 0)) {
        case ACT_NONE * 2: return conv_igemm_tcgen05<ACT_NONE, false, false, false>;
        case ACT_NONE * 2 + 1: return conv_igemm_tcgen05<ACT_NONE, true, false, false>;
        case ACT_RELU * 2: return conv_igemm_tcgen05<ACT_RELU, false, false, false>;
        case ACT_RELU * 2 + 1: return conv_igemm_tcgen05<ACT_RELU, true, false, false>;
        case ACT_GELU * 2: return conv_igemm_tcgen05<ACT_GELU, false, false, false>;
        case ACT_GELU * 2 + 1: return conv_igemm_tcgen05<ACT_GELU, true, false, false>;
        case ACT_SIGMOID * 2: return conv_igemm_tcgen05<ACT_SIGMOID, false, false, false>;
        case ACT_SIGMOID * 2 + 1: return conv_igemm_tcgen05<ACT_SIGMOID, true, false, false>;
        case ACT_HSWISH * 2: return conv_igemm_tcgen05<ACT_HSWISH, false, false, false>;
        case ACT_HSWISH * 2 + 1: return conv_igemm_tcgen05<ACT_HSWISH, true, false, false>;
        case ACT_SWISH * 2: return conv_igemm_tcgen05<ACT_SWISH, false, false, false>;
        case ACT_SWISH * 2 + 1: return conv_igemm_tcgen05<ACT_SWISH, true, false, false>;
        default: return nullptr;
    }
}

int launch_conv(Engine* e, const ConvPlan& plan) {
    static DeviceOnce attr_once;
    if (attr_once.need(e->device)) {
        cudaError_t attr_rc = cudaSuccess;
        auto set = [&](IGemmKernel k) {
            if (attr_rc == cudaSuccess && k != nullptr)
                attr_rc = cudaFuncSetAttribute(reinterpret_cast<const void*>(k),
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
        };
        for (int act = 0; act <= ACT_SWISH; ++act)
            for (int f = 0; f < 2; ++f) set(pick_kernel(act, f, 0, 0));
        set(pick_kernel(ACT_NONE, 1, 1, 0));
        set(pick_kernel(ACT_NONE, 1, 0, 1));
        if (attr_rc != cudaSuccess)
            return set_err(e, DV_ERR_CUDA, "cudaFuncSetAttribute(conv_igemm_tcgen05): %s", cudaGetErrorString(attr_rc));
        attr_once.mark(e->device);
    }
    IGemmKernel k = pick_kernel(plan.prm.act, plan.prm.out_f32, plan.res_f32, plan.prm.arg_out != nullptr);
    if (!k)
        return set_err(e, DV_ERR_ARG, "launch %s: unsupported epilogue (act %d, out_f32 %d, res_f32 %d, argmax %d)",
                       plan.name.c_str(), plan.prm.act, plan.prm.out_f32, plan.res_f32, plan.prm.arg_out != nullptr);
    e->launch_begin("conv_igemm_tcgen05", plan.name, plan.flops, plan.bytes);
    k<<<plan.grid, kIGemmThreads, plan.smem, e->stream>>>(plan.prm);
    e->launch_end();
    cudaError_t st = cudaGetLastError();
    if (st != cudaSuccess)
        return set_err(e, DV_ERR_CUDA, "launch %s failed: %s", plan.name.c_str(), cudaGetErrorString(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ conv_win_tcgen05
// Patch mode (default for stride 1; DV_WINPATCH=0 selects the replicated-window mode): see win_conv_params.h.  Verified on
// B200: an un-swizzled K-major descriptor with LBO = 16 B and 16-byte row pitch (overlapping core matrices) reads the
// staged patch as the implicit im2col matrix.
bool win_patch_enabled() {
    static const bool on = !(getenv("DV_WINPATCH") && atoi(getenv("DV_WINPATCH")) == 0);
    return on;
}

int plan_win_conv(Engine* e, const Tensor& in_padded, int stride, int KR, int Ho, int Wo, const __half* w, const float* bias, int Cout,
                  int act, const Tensor& out_padded, int opad, WinConvPlan* plan, const char* name) {
    WinConvParams& p = plan->prm;
    memset(&p, 0, sizeof(p));
    const int cpp = in_padded.C;
    if ((cpp != 8 && cpp != 16) || in_padded.ld != 0 || (stride != 1 && stride != 2) || KR < 1 || KR > 7)
        return set_err(e, DV_ERR_UNSUPPORTED, "%s: window conv needs a dense 8 / 16 channel padded input, stride 1 / 2, <= 7 filter rows", name);
    const int wpx = 64 / cpp;  // pixels per 128-byte window
    if (in_padded.W < stride * (Wo - 1) + wpx || in_padded.H < stride * (Ho - 1) + KR)
        return set_err(e, DV_ERR_ARG, "%s: padded input too small (%dx%d for %dx%d outputs)", name, in_padded.H, in_padded.W, Ho, Wo);
    if (Cout % 8 || Cout > 64 || (act != ACT_NONE && act != ACT_RELU)) return set_err(e, DV_ERR_UNSUPPORTED, "%s: window conv needs Cout %% 8 == 0, <= 64, ReLU / none", name);
    if (out_padded.C != Cout || out_padded.H < Ho + 2 * opad || out_padded.W < Wo + opad || out_padded.N != in_padded.N)
        return set_err(e, DV_ERR_ARG, "%s: bad output tensor", name);
    p.in = in_padded.p;
    p.Hp = in_padded.H;
    p.Wp = in_padded.W;
    p.cpp = cpp;
    p.stride = stride;
    p.KR = KR;
    p.Nimg = in_padded.N;
    p.Ho = Ho;
    p.Wo = Wo;
    p.patch = (win_patch_enabled() && stride == 1) ? 1 : 0;
    if (p.patch) {
        p.TH = 16;
        p.TW = 8;
        p.planes = cpp / 8;
        p.PR = p.TH + KR - 1;
    } else {
        choose_patch(Ho, Wo, &p.TH, &p.TW);
    }
    p.tiles_y = (Ho + p.TH - 1) / p.TH;
    p.tiles_x = (Wo + p.TW - 1) / p.TW;
    p.m_tiles = p.Nimg * p.tiles_x * p.tiles_y;
    p.BLOCK_N = Cout <= 16 ? 16 : Cout <= 32 ? 32 : 64;
    p.Cout = Cout;
    p.act = act;
    p.bias = bias;
    p.out = out_padded.p;
    p.oHp = out_padded.H;
    p.oWp = out_padded.W;
    p.opad = opad;
    p.out_ld = out_padded.ldc();
    const size_t b_bytes = static_cast<size_t>(p.BLOCK_N) * 128;
    const size_t a_stage = p.patch ? static_cast<size_t>(p.planes) * p.PR * 256 : 128 * 128;
    int stages = static_cast<int>((200 * 1024 - KR * b_bytes) / a_stage);
    if (stages > 12) stages = 12;
    if (stages < 2) return set_err(e, DV_ERR_UNSUPPORTED, "%s: no room for the A ring", name);
    p.num_stages = stages;
    plan->smem = KR * b_bytes + static_cast<size_t>(stages) * a_stage + 1024;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(KR) * 64, static_cast<uint64_t>(Cout)};
        uint64_t str[1] = {static_cast<uint64_t>(KR) * 64 * 2};
        uint32_t box[2] = {64, static_cast<uint32_t>(p.BLOCK_N)};
        DV_TRY(encode_map(e, &p.tmB, w, 2, dims, str, box, 128, name));
    }
    plan->grid = p.m_tiles < e->num_sms ? p.m_tiles : e->num_sms;
    plan->name = std::string(name) + (p.patch ? "[patch]" : "");
    const double px = static_cast<double>(p.Nimg) * Ho * Wo;
    plan->flops = 2.0 * px * KR * 64 * Cout;  // as issued (window padding included)
    plan->bytes = 2.0 * static_cast<double>(in_padded.elems()) + 2.0 * px * Cout + 2.0 * KR * 64 * Cout;
    return 0;
}

int launch_win_conv(Engine* e, const WinConvPlan& plan, double algorithmic_flops) {
    static DeviceOnce attr_once;
    if (attr_once.need(e->device)) {
        cudaError_t attr_rc = cudaFuncSetAttribute(reinterpret_cast<const void*>(conv_win_tcgen05<ACT_RELU>), cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
        if (attr_rc == cudaSuccess)
            attr_rc = cudaFuncSetAttribute(reinterpret_cast<const void*>(conv_win_tcgen05<ACT_NONE>), cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
        if (attr_rc != cudaSuccess) return set_err(e, DV_ERR_CUDA, "cudaFuncSetAttribute(conv_win_tcgen05): %s", cudaGetErrorString(attr_rc));
        attr_once.mark(e->device);
    }
    e->launch_begin("conv_win_tcgen05", plan.name, algorithmic_flops > 0 ? algorithmic_flops : plan.flops, plan.bytes);
    if (plan.prm.act == ACT_RELU) conv_win_tcgen05<ACT_RELU><<<plan.grid, kWinThreads, plan.smem, e->stream>>>(plan.prm);
    else conv_win_tcgen05<ACT_NONE><<<plan.grid, kWinThreads, plan.smem, e->stream>>>(plan.prm);
    e->launch_end();
    cudaError_t st = cudaGetLastError();
    if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "launch %s failed: %s", plan.name.c_str(), cudaGetErrorString(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ mlp_fused_tcgen05
bool mlp_fused_supported(int C) {
    const char* s = getenv("DV_MLP_FUSED");  // read at plan time (not cached): the A/B test builds one engine each way
    return !(s && atoi(s) == 0) && (C == 96 || C == 192 || C == 256);
}

template <int C>
static int launch_mlp_c(Engine* e, const MlpPlan& plan) {
    static DeviceOnce attr_once;
    if (attr_once.need(e->device)) {
        cudaError_t attr_rc = cudaFuncSetAttribute(reinterpret_cast<const void*>(mlp_fused_tcgen05<C>),
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, MlpCfg<C>::SMEM);
        if (attr_rc != cudaSuccess)
            return set_err(e, DV_ERR_CUDA, "cudaFuncSetAttribute(mlp_fused_tcgen05<%d>, %d): %s", C, MlpCfg<C>::SMEM,
                           cudaGetErrorString(attr_rc));
        attr_once.mark(e->device);
    }
    e->launch_begin("mlp_fused_tcgen05", plan.name, plan.flops, plan.bytes);
    mlp_fused_tcgen05<C><<<plan.grid, kMlpThreads, MlpCfg<C>::SMEM, e->stream>>>(plan.prm);
    e->launch_end();
    cudaError_t st = cudaGetLastError();
    if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "launch %s failed: %s", plan.name.c_str(), cudaGetErrorString(st));
    return 0;
}

int plan_mlp(Engine* e, const __half* h, int M, int C, const __half* w1, const float* b1, const __half* w2, const float* b2,
             float* x, MlpPlan* plan, const char* name) {
    if (!(C == 96 || C == 192 || C == 256)) return set_err(e, DV_ERR_UNSUPPORTED, "%s: fused MLP needs C in {96, 192, 256}", name);
    if ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(w2) |
         reinterpret_cast<uintptr_t>(x)) & 15)
        return set_err(e, DV_ERR_UNSUPPORTED, "%s: fused MLP operands must be 16-byte aligned", name);
    MlpParams& p = plan->prm;
    memset(&p, 0, sizeof(p));
    const uint64_t um = static_cast<uint64_t>(M), uc = static_cast<uint64_t>(C);
    {
        uint64_t dims[2] = {uc, um}, str[1] = {uc * 2};
        uint32_t box[2] = {64, 128};
        DV_TRY(encode_map(e, &p.tmA, h, 2, dims, str, box, 128, name));
    }
    {
        uint64_t dims[2] = {uc, 4 * uc}, str[1] = {uc * 2};
        uint32_t box[2] = {64, 64};
        DV_TRY(encode_map(e, &p.tmW1, w1, 2, dims, str, box, 128, name));
    }
    {
        uint64_t dims[2] = {4 * uc, uc}, str[1] = {4 * uc * 2};
        uint32_t box[2] = {64, static_cast<uint32_t>(C)};
        DV_TRY(encode_map(e, &p.tmW2, w2, 2, dims, str, box, 128, name));
    }
    p.b1 = b1;
    p.b2 = b2;
    p.x = x;
    p.M = M;
    p.m_tiles = (M + 127) / 128;
    p.dbg = getenv("DV_MLP_DEBUG") ? atoi(getenv("DV_MLP_DEBUG")) : 0;
    plan->C = C;
    plan->grid = p.m_tiles < e->num_sms ? p.m_tiles : e->num_sms;
    plan->smem = C == 96 ? MlpCfg<96>::SMEM : C == 192 ? MlpCfg<192>::SMEM : MlpCfg<256>::SMEM;
    plan->flops = 2.0 * M * C * 4.0 * C * 2.0;
    plan->bytes = static_cast<double>(M) * C * (2.0 + 4.0 + 4.0) + 16.0 * C * C;  // h read, x read + written, both weights once
    plan->name = name;
    return 0;
}

int launch_mlp(Engine* e, const MlpPlan& plan) {
    switch (plan.C) {
        case 96: return launch_mlp_c<96>(e, plan);
        case 192: return launch_mlp_c<192>(e, plan);
        case 256: return launch_mlp_c<256>(e, plan);
    }
    return set_err(e, DV_ERR_UNSUPPORTED, "launch %s: fused MLP C = %d", plan.name.c_str(), plan.C);
}

// ------------------------------------------------------------------------------------------------ dcn_fused_tcgen05
bool dcn_fused_enabled() {
    const char* s = getenv("DV_DCN_FUSED");  // read at plan time: the A/B test builds one engine each way
    return !(s && atoi(s) == 0);
}

int plan_dcn(Engine* e, const Tensor& in, const float* om, const __half* w, const float* bias, int cout, int act, const Tensor& out,
             DcnPlan* plan, const char* name) {
    if ((in.C % 64) || in.lo > 0 || out.lo > 0 || !(cout == 64 || cout == 128 || cout == 256) || out.C != cout || out.N != in.N ||
        out.H != in.H || out.W != in.W || (in.ldc() % 8) || (out.ldc() % 8))
        return set_err(e, DV_ERR_UNSUPPORTED, "%s: fused DCN needs C %% 64 == 0, cout in {64, 128, 256}, fp16 operands", name);
    if (static_cast<long long>(in.H) * in.W * in.ldc() * 2 >= (1LL << 31))  // 32-bit byte offsets inside an image
        return set_err(e, DV_ERR_UNSUPPORTED, "%s: fused DCN image too large", name);
    DcnParams& p = plan->prm;
    memset(&p, 0, sizeof(p));
    const uint64_t K = 9ull * in.C;
    {
        uint64_t dims[2] = {K, static_cast<uint64_t>(cout)}, str[1] = {K * 2};
        uint32_t box[2] = {64, static_cast<uint32_t>(cout)};
        DV_TRY(encode_map(e, &p.tmB, w, 2, dims, str, box, 128, name));
    }
    p.in = in.p;
    p.om = om;
    p.bias = bias;
    p.out = out.p;
    p.N = in.N;
    p.H = in.H;
    p.W = in.W;
    p.C = in.C;
    p.ldi = in.ldc();
    p.ldo = out.ldc();
    p.cout = cout;
    p.tiles_x = (in.W + kDcnTW - 1) / kDcnTW;
    p.tiles_y = (in.H + kDcnTH - 1) / kDcnTH;
    p.n_tiles = p.tiles_x * p.tiles_y * in.N;
    p.stages = cout == 64 ? 4 : cout == 128 ? 4 : 3;
    if (getenv("DV_DCN_STAGES")) p.stages = std::min(kDcnMaxStages, std::max(2, atoi(getenv("DV_DCN_STAGES"))));
    p.act = act;
    p.dbg = getenv("DV_DCN_DEBUG") ? atoi(getenv("DV_DCN_DEBUG")) : 0;
    plan->grid = p.n_tiles < e->num_sms ? p.n_tiles : e->num_sms;
    plan->smem = dcn_smem_bytes(p.stages, cout);
    if (plan->smem > 220 * 1024) return set_err(e, DV_ERR_UNSUPPORTED, "%s: fused DCN shared memory %zu", name, plan->smem);
    const double px = static_cast<double>(in.N) * in.H * in.W;
    plan->flops = 2.0 * px * static_cast<double>(K) * cout;
    plan->bytes = px * (2.0 * in.C + 128.0 + 2.0 * cout) + 2.0 * static_cast<double>(K) * cout;  // input, offsets / masks, output, weights: once
    plan->name = name;
    return 0;
}

int launch_dcn(Engine* e, const DcnPlan& plan) {
    static DeviceOnce attr_once;
    if (attr_once.need(e->device)) {
        cudaError_t attr_rc = cudaFuncSetAttribute(reinterpret_cast<const void*>(dcn_fused_tcgen05), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   220 * 1024);
        if (attr_rc != cudaSuccess) return set_err(e, DV_ERR_CUDA, "cudaFuncSetAttribute(dcn_fused_tcgen05): %s", cudaGetErrorString(attr_rc));
        attr_once.mark(e->device);
    }
    e->launch_begin("dcn_fused_tcgen05", plan.name, plan.flops, plan.bytes);
    dcn_fused_tcgen05<<<plan.grid, kDcnThreads, plan.smem, e->stream>>>(plan.prm);
    e->launch_end();
    cudaError_t st = cudaGetLastError();
    if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "launch %s failed: %s", plan.name.c_str(), cudaGetErrorString(st));
    return 0;
}

}  // namespace dv
