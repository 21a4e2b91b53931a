/*
 * grpg_b200.h -- C-ABI of the B200-native differentiable 3D-Gaussian rasterizer.
 *
 * This is the drop-in boundary for the hot path of GimpelZhang/GaussianRPG's
 * `diff_gaussian_rasterization` extension.  Every entry point takes plain device
 * pointers, sizes and a cudaStream_t (as void*); no torch type crosses the ABI.
 * Each function cites the reference interface it replaces (paths relative to
 * submodules/diff-gaussian-rasterization/ of the reference).
 *
 * Return value: 0 on success, non-zero on error; grpg_last_error() returns a
 * thread-local, NUL-terminated description of the most recent failure.
 *
 * All device buffers are owned by the caller (the Python binding allocates them
 * as torch tensors, exactly as rasterize_points.cu:70-83 does in the reference).
 * The three workspaces (geometry / binning / image) are opaque byte blobs whose
 * size is obtained from the grpg_*_workspace_bytes() queries; their internal
 * layout is private, but grpg_*_layout() exposes the sub-array offsets that the
 * parity tests need (radii, tiles_touched, point_list, ranges, n_contrib).
 */
#ifndef GRPG_B200_H_
#define GRPG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRPG_NUM_CHANNELS 3   /* cuda_rasterizer/config.h:15 */
#define GRPG_BLOCK_X 16       /* cuda_rasterizer/config.h:17 */
#define GRPG_BLOCK_Y 16       /* cuda_rasterizer/config.h:18 */
#define GRPG_MAX_SEMANTIC_BWD 32 /* reference: NUM_CLASSES 20 (config.h:16); we lift it to 32 and raise beyond */

/* ------------------------------------------------------------------------- */
/* Workspace layout (private, exposed for the parity tests only).            */
/* ------------------------------------------------------------------------- */
typedef struct grpg_geom_layout {
    size_t total_bytes;
    size_t rec;            /* float4[3*P]: (x,y,hx,hy) (A,B,C,opacity) (r,g,b,depth)        */
    size_t depth_key;      /* uint32[P]: float bits of view depth, 0xFFFFFFFF if culled     */
    size_t rect;           /* uint32[2*P]: xmin | xmax<<16, ymin | ymax<<16 of the BINNED rectangle */
    size_t tiles_touched;  /* uint32[P]: tiles of the binned rectangle                      */
    size_t cov3d;          /* float[6*P]                                                    */
    size_t clamped;        /* uint8[P]: bit c set when SH colour channel c was clamped      */
    size_t sorted_idx;     /* uint32[P]: Gaussian ids ordered by (depth bits, id)           */
    size_t offsets;        /* uint32[P]: exclusive scan of tiles_touched in sorted order    */
    size_t scratch;        /* sort ping-pong buffers, histograms, look-back state           */
    size_t tile_mask;      /* uint32[P]: for binned rectangles of <= 32 tiles, bit (ty - y0) * w + (tx - x0) is set when the
                            * tile can hold a pixel with alpha >= 1/255 (only those tiles are binned); all ones otherwise */
    size_t num_rendered;   /* uint64[4] device counters: num_binned, num_rendered (the reference's), overflow flag of
                            * the static-capacity mode, Gaussians in the depth order (those with tiles_touched != 0) */
} grpg_geom_layout;

typedef struct grpg_binning_layout {
    size_t total_bytes;
    size_t point_list;     /* uint32[R]: Gaussian id per instance, sorted by (tile, depth, id) */
    size_t tile_keys;      /* uint32[R]: tile id per sorted instance                           */
    size_t scratch;        /* ping-pong + look-back                                            */
} grpg_binning_layout;

typedef struct grpg_image_layout {
    size_t total_bytes;
    size_t n_contrib;      /* uint32[W*H]  */
    size_t ranges;         /* uint2[tiles] */
} grpg_image_layout;

/* replaces required<GeometryState>(P)  (cuda_rasterizer/rasterizer_impl.h:62-72, rasterizer_impl.cu:155-170) */
int grpg_get_geometry_layout(int P, grpg_geom_layout* out);
/* replaces required<BinningState>(R)   (rasterizer_impl.cu:181-193) */
int grpg_get_binning_layout(long long num_rendered, grpg_binning_layout* out);
/* replaces required<ImageState>(W*H)   (rasterizer_impl.cu:172-179) */
int grpg_get_image_layout(int width, int height, grpg_image_layout* out);

/* ------------------------------------------------------------------------- */
/* Forward.  Replaces CudaRasterizer::Rasterizer::forward                     */
/* (cuda_rasterizer/rasterizer.h:33-64, rasterizer_impl.cu:197-343) as bound  */
/* by RasterizeGaussiansCUDA (rasterize_points.cu:35-124).                    */
/* It is split in two calls because the binning workspace is sized by the     */
/* instance count R, which the caller must allocate (the reference does the   */
/* same through its resize callback after a blocking D2H copy,                */
/* rasterizer_impl.cu:283-288).                                               */
/* ------------------------------------------------------------------------- */
typedef struct grpg_forward_args {
    int P;              /* number of Gaussians                                            */
    int D;              /* active SH degree               (raster_settings.sh_degree)     */
    int M;              /* SH coefficients per Gaussian   (sh.size(1), 0 if no SH)        */
    int S;              /* semantic channels              (semantics.size(1))             */
    int width, height;
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int prefiltered;
    int debug;
    const float* background;     /* [3]      device */
    const float* means3D;        /* [P,3]    device */
    const float* shs;            /* [P,M,3]  device or NULL */
    const float* colors_precomp; /* [P,3]    device or NULL */
    const float* semantics;      /* [P,S]    device or NULL when S==0 */
    const float* opacities;      /* [P,1]    device */
    const float* scales;         /* [P,3]    device or NULL */
    const float* rotations;      /* [P,4]    device or NULL */
    const float* cov3D_precomp;  /* [P,6]    device or NULL */
    const float* viewmatrix;     /* [16]     device, column-major-in-memory as the caller passes it */
    const float* projmatrix;     /* [16]     device */
    const float* cam_pos;        /* [3]      device */
    float* out_color;            /* [3,H,W]  device */
    float* out_depth;            /* [1,H,W]  device */
    float* out_alpha;            /* [1,H,W]  device */
    float* out_semantic;         /* [S,H,W]  device or NULL */
    int*   radii;                /* [P]      device */
    void*  geom_ws;              /* grpg_geom_layout.total_bytes    */
    void*  binning_ws;           /* grpg_binning_layout.total_bytes (stage 2 only) */
    void*  image_ws;             /* grpg_image_layout.total_bytes   */
    void*  stream;               /* cudaStream_t */
    /* Multi-GPU tile-row sharding (no reference counterpart).  With stride k > 1 this call
     * handles only the tile rows r with r % k == phase: instances are emitted, sorted and blended for
     * those rows, and out_color/out_depth/out_alpha/out_semantic/n_contrib use the COMPACT band layout
     * [C, rows_local*16, W] (band row i = tile row i*k + phase).  radii stay global.  stride <= 1 = whole frame. */
    int tile_row_stride;
    int tile_row_phase;
    /* non-zero: no backward will follow (inference): the per-Gaussian cov3D / SH-clamp state that only
     * grpg_backward reads is not written (saves 25 B per visible Gaussian of HBM writes) */
    int forward_only;
    /* Binning rectangle.  0 (default): the reference's rectangle (getRect of the 3-sigma radius,
     * auxiliary.h:46-60) clipped to the tiles that contain a pixel where alpha can reach 1/255 for the stored
     * conic/opacity; everything outside is an instance the reference bins, sorts and then skips in the blend
     * (forward.cu:418-426).  Images, radii and gradients are identical; the instance list is the reference's
     * list minus those dead entries (same relative order).  1: bin exactly the reference's rectangle, so that
     * tiles_touched / point_list / ranges / n_contrib equal the reference's buffers element for element
     * (used by the index-parity tests and for debugging). */
    int reference_binning;
    /* Fused band all-gather over NVLink (multi-GPU, with tile_row_stride > 1): when n_peer_frames > 0 the blend
     * kernel also stores every pixel of its band straight into each peer's FULL frame
     * peer_frames[i] = float[(5 + S), H, W] (channels: 0-2 colour, 3 depth, 4 alpha, 5.. semantics), addressed
     * with the true pixel row.  The buffers must be peer-mapped (e.g. torch symmetric memory); the caller
     * synchronises the ranks around the call. */
    int n_peer_frames;
    float* peer_frames[8];
} grpg_forward_args;

/* number of tile rows owned by (stride, phase) for an image of `height` pixels, and the pixel height of
 * the compact band image (== height when stride <= 1) */
int grpg_band_rows(int height, int stride, int phase);
int grpg_band_height(int height, int stride, int phase);

/* Stage 1: projection (preprocessCUDA forward.cu:155-256), depth ordering and the
 * prefix sum of tile counts (rasterizer_impl.cu:280).  After synchronising `stream` -- the one host sync of
 * the forward path (reference: rasterizer_impl.cu:284) -- writes to the host
 *   *num_binned   = instances this library bins (sizes grpg_get_binning_layout, pass it to grpg_forward_render),
 *   *num_rendered = instances the reference bins = the `num_rendered` RasterizeGaussiansCUDA returns
 *                   (may be NULL; equals *num_binned when reference_binning != 0). */
int grpg_forward_geometry(const grpg_forward_args* a, int* num_binned, int* num_rendered);

/* Stage 2: instance emission (duplicateWithKeys rasterizer_impl.cu:70-111), the
 * (tile | depth | id) ordering (SortPairs rasterizer_impl.cu:306-311), tile ranges
 * (identifyTileRanges :116-138) and the per-tile front-to-back blend (renderCUDA
 * forward.cu:340-467). */
int grpg_forward_render(const grpg_forward_args* a, int num_binned);

/* Both stages in one call when the caller can guess the size of the binning workspace (e.g. from the previous
 * frame): `a->binning_ws` holds `binning_capacity_bytes`.  If the instances fit, stage 2 is launched from inside the
 * call, right behind the host sync, and 0 is returned.  If they do not, stage 1 is complete, GRPG_NEED_BINNING is
 * returned and the caller allocates grpg_get_binning_layout(*num_binned) bytes and calls grpg_forward_render. */
#define GRPG_NEED_BINNING 2
int grpg_forward(const grpg_forward_args* a, size_t binning_capacity_bytes, int* num_binned, int* num_rendered);

/* The whole forward WITHOUT a host synchronisation (no reference counterpart: the reference blocks on a device-to-host
 * copy of num_rendered, rasterizer_impl.cu:284): `a->binning_ws` holds grpg_get_binning_layout(capacity) bytes, every
 * kernel behind the prefix sum reads the instance count from device memory and its grid covers `capacity`.  The call
 * only enqueues work on `a->stream`, so it can be captured into a CUDA graph.  If the frame needs more than `capacity`
 * instances the surplus is dropped (the image is then incomplete) and the overflow counter is set -- the caller must
 * read the counters and re-render with a larger capacity: when `counts_host` (4 x uint64 of pinned host memory) is not
 * NULL, (num_binned, num_rendered, overflow, depth-sorted Gaussians) are copied into it asynchronously on the stream.
 * grpg_backward takes R = capacity for a forward made this way. */
int grpg_forward_static(const grpg_forward_args* a, long long capacity, unsigned long long* counts_host);

/* ------------------------------------------------------------------------- */
/* Backward.  Replaces CudaRasterizer::Rasterizer::backward                   */
/* (rasterizer.h:85-115, rasterizer_impl.cu:396-506) as bound by              */
/* RasterizeGaussiansBackwardCUDA (rasterize_points.cu:126-220).              */
/* All dL_* outputs are fully written by the call (no pre-zeroing needed).    */
/* ------------------------------------------------------------------------- */
typedef struct grpg_backward_args {
    int P, D, M, S, R;
    int width, height;
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int debug;
    const float* background;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* semantics;
    const float* alphas;          /* out_alpha of the forward [1,H,W] */
    const float* scales;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
    const int*   radii;
    const void*  geom_ws;
    const void*  binning_ws;
    const void*  image_ws;
    const float* dL_dpix;         /* [3,H,W] */
    const float* dL_dpix_depth;   /* [1,H,W] */
    const float* dL_dalphas;      /* [1,H,W] */
    const float* dL_dpix_semantic;/* [S,H,W] or NULL */
    /* dL_dmean2D, dL_dconic, dL_dcolor, dL_ddepth, dL_dcov3D, dL_dscale and dL_drot may be NULL: not written then */
    float* dL_dmean2D;            /* [P,3]  (x, y in NDC units, |x|+|y|)  backward.cu:625-628 */
    float* dL_dconic;             /* [P,4]  (.z unused, written 0)        backward.cu:633-635 */
    float* dL_dopacity;           /* [P,1] */
    float* dL_dcolor;             /* [P,3] */
    float* dL_ddepth;             /* [P,1] */
    float* dL_dmean3D;            /* [P,3] */
    float* dL_dcov3D;             /* [P,6] */
    float* dL_dsh;                /* [P,M,3] or NULL when M==0 */
    float* dL_dscale;             /* [P,3] */
    float* dL_drot;               /* [P,4] */
    float* dL_dsemantic;          /* [P,S] or NULL when S==0 */
    void*  grad_ws;               /* [P][12] floats: per-Gaussian 2D gradient record (mean2D xyz, conic xyw, opacity, rgb, depth) */
    void*  stream;
    int tile_row_stride;          /* band of the forward call whose workspaces are passed (see grpg_forward_args); */
    int tile_row_phase;           /* alphas and dL_dpix* then use the compact band layout                          */
    /* stages: 1 = blend backward only (fills grad_ws for all P Gaussians from this band's pixels),
     *         2 = geometry backward only, for Gaussians [p_begin, p_begin + p_count): grad_ws and every dL_*
     *             output then hold p_count rows (row i = Gaussian p_begin + i) while the inputs stay full-size,
     *         3 = both (p_begin = 0, p_count = P).  0 is treated as 3.
     * Between the stages a multi-GPU caller reduce-scatters grad_ws over the Gaussian axis. */
    int stages;
    int p_begin, p_count;
    /* Fused record all-reduce over NVLink (multi-GPU, stage 2): when n_peer_grad > 0 the per-Gaussian backward
     * ignores grad_ws and instead sums, in index order, the full-size [P][12] record buffers peer_grad_ws[0..n)
     * (peer-mapped device pointers, this rank's own buffer included) while it loads them -- the reduction
     * happens in the consumer's load path, there is no separate collective.  Only visible Gaussians are read.
     * The caller orders the ranks: all stage-1 calls complete before any stage-2 call starts, and no buffer
     * is reused before every rank's stage 2 has finished. */
    int n_peer_grad;
    const float* peer_grad_ws[8];
    /* With a band (tile_row_stride > 1): non-zero = dL_dpix, dL_dpix_depth, dL_dalphas, dL_dpix_semantic are FULL
     * frames [C,H,W] indexed by the true pixel row (the caller need not cut the band out of the loss gradient);
     * `alphas` keeps the compact band layout of the forward call either way. */
    int pixel_grads_full_frame;
    int reserved_;
} grpg_backward_args;

size_t grpg_backward_workspace_bytes(int P, int S);
int grpg_backward(const grpg_backward_args* a);

/* replaces CudaRasterizer::Rasterizer::markVisible (rasterizer.h:25-31, rasterizer_impl.cu:141-153) */
int grpg_mark_visible(int P, const float* means3D, const float* viewmatrix,
                      const float* projmatrix, uint8_t* present, void* stream);

/* replaces CudaRasterizer::Rasterizer::visible_filter (rasterizer.h:66-83, rasterizer_impl.cu:345-392) */
int grpg_visible_filter(int P, int width, int height,
                        const float* means3D, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix,
                        float tan_fovx, float tan_fovy,
                        int* radii, float* means2D, void* stream);

/* ------------------------------------------------------------------------- */
/* Multi-GPU gradient-record exchange over NVLink peer memory (no reference    */
/* counterpart).  After the blend backward of a tile-row band, rank  */
/* r holds partial 48-byte records for the Gaussians that touch ITS rows only   */
/* (a quarter of them on 8 GPUs); the per-Gaussian backward of Gaussian g runs   */
/* on the rank that owns g's slice of the Gaussian axis (slice = ceil(P/world)   */
/* consecutive ids).  Instead of reduce-scattering the dense [P,12] buffer      */
/* (96 MB per rank at 2 M Gaussians), grpg_exchange_pack stores the (id, record) */
/* pairs of the in-band Gaussians straight into the owner's inbox through        */
/* peer-mapped pointers (48-byte entries, warp-aggregated slot allocation), and  */
/* after a barrier between the ranks grpg_exchange_accumulate adds the entries   */
/* of the local inbox into the local record buffer.  Worst-case capacity (every  */
/* Gaussian of a slice from every peer) is reserved, so there is no overflow.    */
/* ------------------------------------------------------------------------- */
typedef struct grpg_exchange_args {
    int P, world, rank;
    const void* geom_ws;     /* this rank's forward geometry workspace (which Gaussians are in the band)          */
    float* grad_rec;         /* [>= P, 12] device: partial records in, (own slice) complete records out            */
    void* inbox[8];          /* inbox[q]: peer-mapped base of rank q's inbox (inbox[rank] is local memory)         */
    void* stream;
} grpg_exchange_args;

size_t grpg_exchange_inbox_bytes(int P, int world);
/* stores this rank's in-band records of foreign slices into their owners' inboxes and publishes the counts */
int grpg_exchange_pack(const grpg_exchange_args* a);
/* to be called after every rank's grpg_exchange_pack has completed (caller's barrier): adds the local inbox's entries
 * into grad_rec; afterwards rows [rank*slice, (rank+1)*slice) of grad_rec hold the sums over all ranks */
int grpg_exchange_accumulate(const grpg_exchange_args* a);

/* Reconstructs the reference's sorted 64-bit keys ((tile << 32) | depth bits,
 * rasterizer_impl.cu:102-104) from our workspaces -- test hook for the bit-exact
 * key comparison. */
int grpg_debug_reference_keys(int P, long long num_rendered, const void* geom_ws,
                              const void* binning_ws, uint64_t* keys_out, void* stream);

/* Test hook for the packed-FP32 blend kernels: evaluates the packed expf restatement (csrc/grpg_common.cuh
 * `expf2_exact`) and libdevice's expf() on every float bit pattern in [first_bits, last_bits] (as positive-sign or
 * negative-sign floats per `negative`) and counts the inputs on which the two differ bitwise (NaN results compare
 * equal); out[0] = mismatches, out[1] = 1 if rcp.approx.ftz(1.0f) == 1.0f exactly (the masking of the packed backward
 * relies on it), out[2] = inputs evaluated.  `out` = 3 device uint64. */
int grpg_debug_packed_math_check(uint32_t first_bits, uint32_t last_bits, int negative,
                                 unsigned long long* out, void* stream);

/* Optional per-kernel timing: between begin and end every kernel this library launches is
 * bracketed by CUDA events on its launching stream.  grpg_profile_end synchronises and writes
 * "name:launches:total_ms" lines into buf.  Used by bench.py for the roofline object only. */
int grpg_profile_begin(void);
int grpg_profile_end(char* buf, size_t buf_len);

const char* grpg_last_error(void);
int grpg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GRPG_B200_H_ */
