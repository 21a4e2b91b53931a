// api.cu -- the C-ABI of libgrpg_b200.so (declared in include/grpg_b200.h).
// Host orchestration only: workspace carving, kernel sequencing, error reporting.
#include <cstdio>
#include <cstring>
#include <string>
#include "grpg_common.cuh"

namespace grpg {
// preprocess_fwd.cu
void launch_preprocess_fwd(const grpg_forward_args* a, float focal_x, float focal_y, uint32_t grid_x, uint32_t grid_y,
                           Rec* rec, uint32_t* depth_key, uint2* rect, uint32_t* tiles_touched, float* cov3d,
                           uint8_t* clamped, uint32_t* tile_mask, unsigned long long* ref_instances, cudaStream_t stream);
void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream);
void launch_visible_filter(int P, int W, int H, const float* means3D, const float* scales, float scale_modifier,
                           const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                           const float* projmatrix, float tan_fovx, float tan_fovy, int* radii, float* means2D,
                           cudaStream_t stream);
// binning.cu
size_t geom_scratch_bytes(int P);
size_t binning_scratch_bytes(long long R);
unsigned long long* prepare_geometry_scratch(int P, void* scratch, cudaStream_t stream);
void run_depth_order_and_scan(int P, uint32_t* depth_key, uint32_t* sorted_idx, const uint32_t* tiles_touched,
                              uint32_t* offsets, void* scratch, unsigned long long* counts, unsigned long long capacity,
                              int num_sms, cudaStream_t stream);
bool exact_tile_cull();  // preprocess_fwd.cu: per-tile masks are written (and read by the emission) only in that mode
void run_instance_binning(int P, long long R, uint32_t grid_x, uint32_t num_tiles, const uint32_t* sorted_idx,
                          const uint32_t* offsets, const uint2* rect, const uint32_t* tile_mask, const void* geom_scratch,
                          uint32_t* tile_keys, uint32_t* point_list, void* scratch, uint2* ranges,
                          const unsigned long long* counts, int static_capacity, int num_sms, cudaStream_t stream);
void launch_reference_keys(long long R, const uint32_t* point_list, const uint32_t* tile_keys, const Rec* rec,
                           unsigned long long* keys, cudaStream_t stream);
void launch_packed_math_check(uint32_t first_bits, uint32_t last_bits, int negative, unsigned long long* out,
                              cudaStream_t stream);
// blend_fwd.cu / blend_bwd.cu / preprocess_bwd.cu
void launch_blend_fwd(const grpg_forward_args* a, const uint2* ranges, const uint32_t* point_list, const Rec* rec,
                      uint32_t* n_contrib, long long num_instances, cudaStream_t stream);
void launch_blend_bwd(const grpg_backward_args* a, const uint2* ranges, const uint32_t* point_list, const Rec* rec,
                      const uint32_t* n_contrib, float* grad_rec, cudaStream_t stream);
void launch_preprocess_bwd(const grpg_backward_args* a, const float* cov3D, const uint8_t* clamped,
                           const float* grad_rec, cudaStream_t stream);
}  // namespace grpg

using namespace grpg;

// ---- optional per-kernel timing (CUDA events on the launching stream) ------------------------
#include <map>
#include <vector>
namespace {
struct ProfEntry { std::string name; cudaEvent_t e0, e1; };
bool g_prof_on = false;
std::vector<ProfEntry> g_prof;
}  // namespace
namespace grpg {
void prof_range_begin(const char* name, cudaStream_t stream) {
    if (!g_prof_on) return;
    ProfEntry e;
    e.name = name;
    cudaEventCreate(&e.e0);
    cudaEventCreate(&e.e1);
    cudaEventRecord(e.e0, stream);
    g_prof.push_back(e);
}
void prof_range_end(cudaStream_t stream) {
    if (!g_prof_on || g_prof.empty()) return;
    cudaEventRecord(g_prof.back().e1, stream);
}
}  // namespace grpg

static thread_local std::string g_last_error;
static int fail(const std::string& msg) {
    g_last_error = msg;
    return 1;
}
static int check_cuda(const char* what, bool sync, cudaStream_t stream) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && sync) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return fail(std::string("[CUDA ERROR] in ") + what + ": " + cudaGetErrorString(e));
    return 0;
}
static int device_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}
// one pinned word per host thread for the single D2H read of the forward
static unsigned long long* pinned_word() {
    static thread_local unsigned long long* p = nullptr;
    if (!p) cudaHostAlloc((void**)&p, 64, cudaHostAllocDefault);
    return p;
}

extern "C" int grpg_loss_fail(const char* msg) { return fail(msg ? msg : "error"); }  // used by loss_ssim.cu

extern "C" {

const char* grpg_last_error(void) { return g_last_error.c_str(); }

int grpg_profile_begin(void) {
    for (auto& e : g_prof) { cudaEventDestroy(e.e0); cudaEventDestroy(e.e1); }
    g_prof.clear();
    g_prof_on = true;
    return 0;
}

// Writes "name:count:total_ms\n" lines into buf; returns the number of distinct kernels.
int grpg_profile_end(char* buf, size_t buf_len) {
    g_prof_on = false;
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<int, double>> acc;
    std::vector<std::string> order;
    for (auto& e : g_prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.e0, e.e1);
        if (!acc.count(e.name)) order.push_back(e.name);
        acc[e.name].first += 1;
        acc[e.name].second += ms;
        cudaEventDestroy(e.e0);
        cudaEventDestroy(e.e1);
    }
    g_prof.clear();
    std::string out;
    for (auto& n : order) {
        char line[256];
        snprintf(line, sizeof line, "%s:%d:%.6f\n", n.c_str(), acc[n].first, acc[n].second);
        out += line;
    }
    if (buf && buf_len) {
        strncpy(buf, out.c_str(), buf_len - 1);
        buf[buf_len - 1] = 0;
    }
    return (int)order.size();
}
int grpg_version(void) { return 100; }

int grpg_band_rows(int height, int stride, int phase) { return band_rows(height, stride, phase); }
int grpg_band_height(int height, int stride, int phase) { return band_height(height, stride, phase); }

int grpg_get_geometry_layout(int P, grpg_geom_layout* out) {
    if (!out || P < 0) return fail("grpg_get_geometry_layout: bad arguments");
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return at; };
    const size_t p = (size_t)P;
    out->rec = take(p * sizeof(Rec));
    out->depth_key = take(p * 4);
    out->rect = take(p * 8);
    out->tiles_touched = take(p * 4);
    out->cov3d = take(p * 24);
    out->clamped = take(p);
    out->sorted_idx = take(p * 4);
    out->offsets = take(p * 4);
    out->scratch = take(geom_scratch_bytes(P));
    out->tile_mask = take(p * 4);
    out->num_rendered = take(64);
    out->total_bytes = o;
    return 0;
}

int grpg_get_binning_layout(long long R, grpg_binning_layout* out) {
    if (!out || R < 0) return fail("grpg_get_binning_layout: bad arguments");
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return at; };
    out->point_list = take((size_t)R * 4);
    out->tile_keys = take((size_t)R * 4);
    out->scratch = take(binning_scratch_bytes(R));
    out->total_bytes = o;
    return 0;
}

int grpg_get_image_layout(int width, int height, grpg_image_layout* out) {
    if (!out || width < 0 || height < 0) return fail("grpg_get_image_layout: bad arguments");
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return at; };
    const size_t tiles = (size_t)((width + 15) / 16) * ((height + 15) / 16);
    out->n_contrib = take((size_t)width * height * 4);
    out->ranges = take(tiles * 8);
    out->total_bytes = o;
    return 0;
}

static int validate_forward(const grpg_forward_args* a) {
    if (!a) return fail("null arguments");
    if (a->P < 0 || a->width <= 0 || a->height <= 0) return fail("bad sizes");
    if (a->P == 0) return 0;
    if (!a->means3D || !a->opacities || !a->viewmatrix || !a->projmatrix || !a->cam_pos || !a->background)
        return fail("missing required input pointer");
    if ((a->shs == nullptr) == (a->colors_precomp == nullptr))
        return fail("Please provide excatly one of either SHs or precomputed colors!");
    if (((a->scales == nullptr || a->rotations == nullptr) && a->cov3D_precomp == nullptr) ||
        ((a->scales != nullptr || a->rotations != nullptr) && a->cov3D_precomp != nullptr))
        return fail("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    if (GRPG_NUM_CHANNELS != 3 && a->colors_precomp == nullptr)
        return fail("For non-RGB, provide precomputed Gaussian colors!");  // rasterizer_impl.cu:245-248
    if (a->shs && (a->D < 0 || a->D > 3 || (a->D + 1) * (a->D + 1) > a->M))
        return fail("sh_degree needs (D+1)^2 <= M coefficients and D <= 3");
    if (a->S < 0 || (a->S > 0 && (!a->semantics || !a->out_semantic))) return fail("bad semantics arguments");
    if ((a->width + 15) / 16 > 65535 || (a->height + 15) / 16 > 65535) return fail("image too large (tile grid > 65535)");
    if (!a->geom_ws || !a->image_ws || !a->radii) return fail("missing workspace");
    if (a->tile_row_stride > 1 && (a->tile_row_phase < 0 || a->tile_row_phase >= a->tile_row_stride))
        return fail("tile_row_phase must lie in [0, tile_row_stride)");
    return 0;
}

// launches of stage 1 (no synchronisation)
static int launch_forward_geometry(const grpg_forward_args* a, unsigned long long capacity) {
    cudaStream_t stream = (cudaStream_t)a->stream;
    grpg_geom_layout L;
    grpg_get_geometry_layout(a->P, &L);
    char* g = (char*)a->geom_ws;
    const float focal_y = a->height / (2.0f * a->tan_fovy);
    const float focal_x = a->width / (2.0f * a->tan_fovx);
    const uint32_t gx = (a->width + 15) / 16, gy = (a->height + 15) / 16;
    unsigned long long* ref_instances = prepare_geometry_scratch(a->P, g + L.scratch, stream);
    launch_preprocess_fwd(a, focal_x, focal_y, gx, gy, (Rec*)(g + L.rec), (uint32_t*)(g + L.depth_key),
                          (uint2*)(g + L.rect), (uint32_t*)(g + L.tiles_touched), (float*)(g + L.cov3d),
                          (uint8_t*)(g + L.clamped), (uint32_t*)(g + L.tile_mask), ref_instances, stream);
    if (a->debug) if (int rc = check_cuda("preprocess", true, stream)) return rc;
    run_depth_order_and_scan(a->P, (uint32_t*)(g + L.depth_key), (uint32_t*)(g + L.sorted_idx),
                             (const uint32_t*)(g + L.tiles_touched), (uint32_t*)(g + L.offsets), g + L.scratch,
                             (unsigned long long*)(g + L.num_rendered), capacity, device_sm_count(), stream);
    return 0;
}

int grpg_forward_geometry(const grpg_forward_args* a, int* num_binned, int* num_rendered) {
    if (!num_binned) return fail("null num_binned");
    *num_binned = 0;
    if (num_rendered) *num_rendered = 0;
    if (int rc = validate_forward(a)) return rc;
    if (a->P == 0) return 0;
    cudaStream_t stream = (cudaStream_t)a->stream;
    if (int rc = launch_forward_geometry(a, 0)) return rc;
    grpg_geom_layout L;
    grpg_get_geometry_layout(a->P, &L);
    unsigned long long* host = pinned_word();
    cudaMemcpyAsync(host, (char*)a->geom_ws + L.num_rendered, 16, cudaMemcpyDeviceToHost, stream);
    if (int rc = check_cuda("forward_geometry", true, stream)) return rc;
    if (host[0] >= (1ull << 31) || host[1] >= (1ull << 31)) return fail("too many Gaussian/tile instances (>= 2^31, the range of the reference's int num_rendered)");
    *num_binned = (int)host[0];
    if (num_rendered) *num_rendered = (int)host[1];
    return 0;
}

static int forward_render_impl(const grpg_forward_args* a, long long num_rendered, int static_capacity) {
    cudaStream_t stream = (cudaStream_t)a->stream;
    const int stride = a->tile_row_stride > 1 ? a->tile_row_stride : 1, phase = a->tile_row_stride > 1 ? a->tile_row_phase : 0;
    grpg_image_layout IL;
    grpg_get_image_layout(a->width, band_height(a->height, stride, phase), &IL);
    char* im = (char*)a->image_ws;
    const uint32_t gx = (a->width + 15) / 16, gy = (uint32_t)band_rows(a->height, stride, phase);
    if (a->P == 0) return 0;
    if (num_rendered > 0 && !a->binning_ws) return fail("missing binning workspace");
    grpg_geom_layout L;
    grpg_get_geometry_layout(a->P, &L);
    grpg_binning_layout BL;
    grpg_get_binning_layout(num_rendered, &BL);
    char* g = (char*)a->geom_ws;
    char* b = (char*)a->binning_ws;
    run_instance_binning(a->P, num_rendered, gx, gx * gy, (const uint32_t*)(g + L.sorted_idx),
                         (const uint32_t*)(g + L.offsets), (const uint2*)(g + L.rect),
                         (!a->reference_binning && exact_tile_cull()) ? (const uint32_t*)(g + L.tile_mask) : nullptr,
                         g + L.scratch,
                         b ? (uint32_t*)(b + BL.tile_keys) : nullptr, b ? (uint32_t*)(b + BL.point_list) : nullptr,
                         b ? b + BL.scratch : nullptr, (uint2*)(im + IL.ranges),
                         (const unsigned long long*)(g + L.num_rendered), static_capacity, device_sm_count(), stream);
    if (a->debug) if (int rc = check_cuda("binning", true, stream)) return rc;
    launch_blend_fwd(a, (const uint2*)(im + IL.ranges), b ? (const uint32_t*)(b + BL.point_list) : nullptr,
                     (const Rec*)(g + L.rec), (uint32_t*)(im + IL.n_contrib), static_capacity ? 0 : num_rendered, stream);
    return check_cuda("forward_render", a->debug != 0, stream);
}

int grpg_forward_render(const grpg_forward_args* a, int num_rendered) {
    if (int rc = validate_forward(a)) return rc;
    return forward_render_impl(a, num_rendered, 0);
}

int grpg_forward(const grpg_forward_args* a, size_t binning_capacity_bytes, int* num_binned, int* num_rendered) {
    if (int rc = grpg_forward_geometry(a, num_binned, num_rendered)) return rc;
    grpg_binning_layout BL;
    grpg_get_binning_layout(*num_binned, &BL);
    if (*num_binned > 0 && (!a->binning_ws || BL.total_bytes > binning_capacity_bytes)) return GRPG_NEED_BINNING;
    return grpg_forward_render(a, *num_binned);
}

int grpg_forward_static(const grpg_forward_args* a, long long capacity, unsigned long long* counts_host) {
    if (int rc = validate_forward(a)) return rc;
    if (capacity <= 0 || capacity >= (1ll << 31)) return fail("grpg_forward_static: capacity must lie in [1, 2^31)");
    if (a->debug) return fail("grpg_forward_static: debug mode synchronises; use grpg_forward");
    if (a->P == 0) return 0;
    if (!a->binning_ws) return fail("missing binning workspace");
    cudaStream_t stream = (cudaStream_t)a->stream;
    if (int rc = launch_forward_geometry(a, (unsigned long long)capacity)) return rc;
    if (int rc = forward_render_impl(a, capacity, 1)) return rc;
    if (counts_host) {
        grpg_geom_layout L;
        grpg_get_geometry_layout(a->P, &L);
        cudaMemcpyAsync(counts_host, (const char*)a->geom_ws + L.num_rendered, 32, cudaMemcpyDeviceToHost, stream);
    }
    cudaError_t e = cudaGetLastError();  // launch-configuration errors only; nothing here waits for the device
    if (e != cudaSuccess) return fail(std::string("[CUDA ERROR] in forward_static: ") + cudaGetErrorString(e));
    return 0;
}

size_t grpg_backward_workspace_bytes(int P, int S) {
    (void)S;
    return align_up((size_t)P * 12 * sizeof(float), 256);
}

int grpg_backward(const grpg_backward_args* a_in) {
    if (!a_in) return fail("null arguments");
    if (a_in->P == 0) return 0;
    grpg_backward_args args = *a_in;
    grpg_backward_args* a = &args;
    if (a->stages == 0) a->stages = 3;
    if (a->stages == 3) { a->p_begin = 0; a->p_count = a->P; }
    if (a->S > GRPG_MAX_SEMANTIC_BWD)
        return fail("backward supports at most 32 semantic channels (reference NUM_CLASSES is 20, config.h:16)");
    if (!a->grad_ws || !a->geom_ws) return fail("missing workspace");
    if (a->p_begin < 0 || a->p_count < 0 || a->p_begin + a->p_count > a->P) return fail("bad Gaussian slice");
    cudaStream_t stream = (cudaStream_t)a->stream;
    const int stride = a->tile_row_stride > 1 ? a->tile_row_stride : 1, phase = a->tile_row_stride > 1 ? a->tile_row_phase : 0;
    grpg_geom_layout L;
    grpg_get_geometry_layout(a->P, &L);
    const char* g = (const char*)a->geom_ws;
    float* grad_rec = (float*)a->grad_ws;
    if (a->stages & 1) {
        if (!a->image_ws) return fail("missing image workspace");
        grpg_binning_layout BL;
        grpg_get_binning_layout(a->R, &BL);
        grpg_image_layout IL;
        grpg_get_image_layout(a->width, band_height(a->height, stride, phase), &IL);
        const char* b = (const char*)a->binning_ws;
        const char* im = (const char*)a->image_ws;
        cudaMemsetAsync(grad_rec, 0, (size_t)a->P * 12 * sizeof(float), stream);
        if (a->S > 0) cudaMemsetAsync(a->dL_dsemantic, 0, (size_t)a->P * a->S * sizeof(float), stream);
        if (a->R > 0) {
            if (!b) return fail("missing binning workspace");  // allocate >= 16 bytes even when nothing was binned
            launch_blend_bwd(a, (const uint2*)(im + IL.ranges), (const uint32_t*)(b + BL.point_list),
                             (const Rec*)(g + L.rec), (const uint32_t*)(im + IL.n_contrib), grad_rec, stream);
            if (a->debug) if (int rc = check_cuda("blend backward", true, stream)) return rc;
        }
    }
    if (a->stages & 2) {
        const float* cov3D = a->cov3D_precomp ? a->cov3D_precomp : (const float*)(g + L.cov3d);
        launch_preprocess_bwd(a, cov3D, (const uint8_t*)(g + L.clamped), grad_rec, stream);
    }
    return check_cuda("backward", a->debug != 0, stream);
}

int grpg_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                      void* stream) {
    (void)projmatrix;
    if (P <= 0) return 0;
    if (!means3D || !viewmatrix || !present) return fail("grpg_mark_visible: null pointer");
    launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
    return check_cuda("mark_visible", false, (cudaStream_t)stream);
}

int grpg_visible_filter(int P, int width, int height, const float* means3D, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                        const float* projmatrix, float tan_fovx, float tan_fovy, int* radii, float* means2D,
                        void* stream) {
    if (P <= 0) return 0;
    if (!means3D || !viewmatrix || !projmatrix || !radii || !means2D) return fail("grpg_visible_filter: null pointer");
    if (!cov3D_precomp && (!scales || !rotations)) return fail("grpg_visible_filter: need scales+rotations or cov3D_precomp");
    launch_visible_filter(P, width, height, means3D, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
                          projmatrix, tan_fovx, tan_fovy, radii, means2D, (cudaStream_t)stream);
    return check_cuda("visible_filter", false, (cudaStream_t)stream);
}

int grpg_debug_reference_keys(int P, long long R, const void* geom_ws, const void* binning_ws, uint64_t* keys_out,
                              void* stream) {
    if (R <= 0) return 0;
    grpg_geom_layout L;
    grpg_get_geometry_layout(P, &L);
    grpg_binning_layout BL;
    grpg_get_binning_layout(R, &BL);
    const char* g = (const char*)geom_ws;
    const char* b = (const char*)binning_ws;
    launch_reference_keys(R, (const uint32_t*)(b + BL.point_list), (const uint32_t*)(b + BL.tile_keys),
                          (const Rec*)(g + L.rec), (unsigned long long*)keys_out, (cudaStream_t)stream);
    return check_cuda("reference_keys", false, (cudaStream_t)stream);
}

int grpg_debug_packed_math_check(uint32_t first_bits, uint32_t last_bits, int negative, unsigned long long* out,
                                 void* stream) {
    if (!out || last_bits < first_bits) return fail("grpg_debug_packed_math_check: bad arguments");
    launch_packed_math_check(first_bits, last_bits, negative, out, (cudaStream_t)stream);
    return check_cuda("packed_math_check", false, (cudaStream_t)stream);
}

}  // extern "C"
