/*
 * fftconv_b200.h — C ABI of the B200-native FFT-convolution hot path.
 *
 * This is the drop-in boundary for the convolution path of klae01/fft-conv-pytorch.
 * The reference has no FFI (it is pure Python on torch); the entry points below are what
 * a binding for that path would call, one per step of the reference's algorithm:
 *
 *   fc_plan_create        argument normalisation + shape algebra of
 *                         fft_conv            (reference fft_conv_pytorch/functional.py:44-66, 76-82)
 *                         fft_conv_transpose  (reference fft_conv_pytorch/functional.py:103-154, 163-169)
 *                         to_ntuple           (reference fft_conv_pytorch/utils.py:4-20; done host-side)
 *   fc_signal_spectrum    F.pad + zero-stuffing + rfftn(signal)   (functional.py:60-62, 126-139, 70, 157)
 *   fc_kernel_spectrum    dilation scatter / transposed regroup + rfftn(kernel).conj()
 *                                                                 (functional.py:49-57, 109-124, 71, 158)
 *   fc_contract           complex_matmul                          (functional.py:11-16)
 *   fc_inverse            irfftn + crop/stride slice + bias add   (functional.py:68-87, 155-174)
 *   fc_conv               the whole call with a cached kernel spectrum (functional.py:19-89 / 92-176)
 *   fc_conv_host          same, host buffers in / host buffers out (what a CPU-tensor caller of the
 *                         reference's fft_conv sees); copies are issued on the given stream
 *
 * Conventions
 *   - All tensors are dense, contiguous, fp32 (complex = interleaved float pairs), NC[D][H]W order as in torch.
 *   - Pointers named d_* are DEVICE pointers, h_* are HOST pointers. The library never allocates or frees
 *     device memory and never synchronises the device; the caller owns the output, the spectrum cache, the
 *     constant table and the workspace (sizes from the fc_*_bytes queries). Work is queued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - Every entry point returns 0 on success, a negative FC_E* code for an invalid/unsupported argument, or a
 *     positive cudaError_t. fc_last_error() returns a thread-local message for the last failure.
 *   - A plan is an immutable host object; it may be shared between threads and streams as long as the
 *     buffers passed with it differ.
 */
#ifndef FFTCONV_B200_H
#define FFTCONV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FC_MAX_ND 3

enum {
  FC_OK = 0,
  FC_EINVAL = -1,       /* malformed problem (the reference would raise or silently return garbage) */
  FC_EUNSUPPORTED = -2, /* well-formed but outside what the kernels cover (e.g. transform extent too large) */
  FC_ENULL = -3         /* required pointer is NULL */
};

/* padding_mode of fft_conv (functional.py:27): "constant" is what nn.py:12 maps torch's "zeros" to. */
enum { FC_PAD_CONSTANT = 0, FC_PAD_REFLECT = 1, FC_PAD_REPLICATE = 2, FC_PAD_CIRCULAR = 3 };

/* One convolution call. Spatial arrays are indexed in torch order (slowest axis first); entries >= ndim are ignored. */
typedef struct fc_problem {
  int32_t ndim;       /* spatial dims, 1..3  (signal.ndim - 2, functional.py:44) */
  int32_t transposed; /* 0: fft_conv, 1: fft_conv_transpose */
  int32_t batch;      /* B */
  int32_t cin;        /* signal channels */
  int32_t cout;       /* output channels */
  int32_t groups;
  int32_t in_size[FC_MAX_ND];     /* signal spatial extents */
  int32_t kernel_size[FC_MAX_ND]; /* kernel spatial extents */
  int32_t stride[FC_MAX_ND];
  int32_t padding[FC_MAX_ND];
  int32_t dilation[FC_MAX_ND];
  int32_t output_padding[FC_MAX_ND]; /* transposed only */
  int32_t padding_mode;              /* FC_PAD_*; forward only */
  int32_t threads;                   /* CTA size override for the pass kernels; 0 = default */
  int32_t flags;                     /* FC_FLAG_* */
  int32_t reserved;
} fc_problem;

enum {
  FC_FLAG_NO_FUSED = 1,     /* generic kernels only: K1 / axis passes / K3 / K4, nothing specialised or fused */
  FC_FLAG_NO_POLYPHASE = 2, /* keep stride/dilation lattices dense (no gcd reduction) */
  FC_FLAG_NO_FAST_R2C = 4,  /* the three below switch the specialised kernels off one by one (tests, A/B timing) */
  FC_FLAG_NO_FAST_C2R = 8,
  FC_FLAG_NO_FUSED_MID = 16,
  FC_FLAG_NO_TC = 32,       /* keep the contraction on the SIMT kernel even when the tensor-core path qualifies */
  FC_FLAG_NO_FAST_C2C = 64, /* keep contiguous complex axis passes on the generic block-level kernel */
  FC_FLAG_NO_SEGMENT = 128, /* never split the first axis of a 2-d problem into overlap-save segments */
  FC_FLAG_NO_PAIR = 256,    /* keep the fused 2-d program on the one-line-per-item kernels (no packed batch pairs) */
  FC_FLAG_PAIR = 512,       /* run it on the packed batch-pair kernels wherever they apply (default: where they measured faster) */
  FC_FLAG_NO_YSTAGE = 1024, /* pair program: keep the whole transform of the fused axis inside the fused kernel */
  FC_FLAG_NO_ROW_FILL = 2048, /* transposed row lattices: leave the bias-only rows to the last kernel (A/B timing) */
  FC_FLAG_NO_STREAM = 4096,  /* keep K4 on the register-path kernel (no tensor-map streaming, fc_stream.cuh; tests, A/B timing) */
  FC_FLAG_SEGMENT = 16384,   /* 1-d fft_conv: run as batch segments (windows of the line) whenever a window length applies, even if the cost model sees no gain; with FC_FLAG_NO_SEGMENT the two pin the (batch-dependent) choice for plans that share a kernel spectrum */
  FC_FLAG_NO_SHORT_SPLIT = 32768, /* 1-d lines of 2048 ... 8192 points: keep the one-pass layout on the generic kernels (default: four-step 64 x N2 on the column / warp-engine kernels) */
  FC_FLAG_STREAM_R2C = 8192  /* run K1 on the bulk-copy / tensor-map-store kernel too (measured equal to the register path at BASELINE c2) */
};

typedef struct fc_plan fc_plan; /* opaque */

/* Shape/size facts of a plan, for the host side (allocation, roofline accounting, tests). */
typedef struct fc_plan_info {
  int32_t ndim;
  int32_t out_size[FC_MAX_ND];   /* output spatial extents (reference functional.py:79 / 144-154) */
  int32_t fft_size[FC_MAX_ND];   /* transform extent N* actually used per axis */
  int32_t n_launches;            /* kernels queued by one fc_conv call */
  int32_t n_launches_kspec;      /* kernels queued by one fc_kernel_spectrum call */
  int32_t fused;                 /* 1 if fc_conv runs the fused axis+contract+axis kernel */
  int32_t segments;              /* overlap-save segments on the first axis of a 2-d problem (1 = unsegmented); fft_size[0]
                                    is then the segment transform length. Segmented plans run through fc_conv /
                                    fc_kernel_spectrum only: the stage calls return FC_EUNSUPPORTED */
  int64_t bins;                  /* half-spectrum bins per (batch, channel) */
  int64_t out_elems;             /* B*Cout*prod(out_size) */
  int64_t xspec_bytes;           /* signal spectrum buffer */
  int64_t kspec_bytes;           /* kernel spectrum buffer (the cacheable object) */
  int64_t yspec_bytes;           /* product spectrum buffer */
  int64_t workspace_bytes;       /* scratch needed by fc_conv / the stage calls (includes x/y spectra) */
  int64_t const_bytes;           /* twiddle table to be initialised once by fc_plan_init_const */
  int64_t algo_bytes_s1, algo_bytes_s2, algo_bytes_s3, algo_bytes_s4; /* SURVEY §8d S1..S4 at N* */
  int64_t kspec_workspace_bytes; /* scratch needed by fc_kernel_spectrum (>= workspace_bytes when the tensor-core layout is built) */
  int32_t tensor_core;           /* 1 if the contraction of fc_conv runs on the tensor cores (fc_tc_* path) */
  int32_t reserved2;
} fc_plan_info;

const char* fc_last_error(void);
const char* fc_version(void);

int fc_plan_create(fc_plan** out, const fc_problem* problem);
void fc_plan_destroy(fc_plan* plan);
int fc_plan_get_info(const fc_plan* plan, fc_plan_info* info);
/* Human-readable plan dump (passes, strides, maps) for tests/debugging; returns bytes written (excl. NUL). */
int fc_plan_describe(const fc_plan* plan, char* buf, size_t buflen);

/* One-time: fill the plan's constant table (twiddle factors) in caller-owned device memory of const_bytes. */
int fc_plan_init_const(const fc_plan* plan, void* d_const, void* stream);

/* Stage calls (unfused pipeline). d_ws: workspace_bytes scratch. The kernel spectrum is a plan-specific object (layout
 * chosen for the plan's contraction kernel): fc_contract takes the channel-major layout [o*Cin/g + i][bins] only and
 * returns FC_EUNSUPPORTED for plans with info.fused / info.tensor_core set; all three stage calls return FC_EUNSUPPORTED
 * for plans with info.segments > 1. Plans made with FC_FLAG_NO_FUSED_MID | FC_FLAG_NO_SEGMENT | FC_FLAG_NO_TC run them. */
int fc_signal_spectrum(const fc_plan* plan, const void* d_const, const float* d_x, float* d_xspec, void* d_ws, void* stream);
int fc_kernel_spectrum(const fc_plan* plan, const void* d_const, const float* d_w, float* d_kspec, void* d_ws, void* stream);
int fc_contract(const fc_plan* plan, const float* d_xspec, const float* d_kspec, float* d_yspec, void* stream);
int fc_inverse(const fc_plan* plan, const void* d_const, const float* d_yspec, const float* d_bias /*nullable*/, float* d_y, void* d_ws, void* stream);

/* Whole call with a cached kernel spectrum: y = conv(x; kspec) + bias. */
int fc_conv(const fc_plan* plan, const void* d_const, const float* d_x, const float* d_kspec, const float* d_bias /*nullable*/,
            float* d_y, void* d_ws, void* stream);

/* Host-buffer variant: copies h_x to d_x_stage, runs fc_conv, copies the result to h_y, all on `stream`
 * (asynchronous when the host buffers are pinned). d_x_stage / d_y_stage are caller-owned device buffers. */
int fc_conv_host(const fc_plan* plan, const void* d_const, const float* h_x, float* d_x_stage, const float* d_kspec,
                 const float* d_bias /*nullable*/, float* d_y_stage, float* h_y, void* d_ws, void* stream);

/* Profiling helpers (bench.py roofline accounting). fc_conv_profiled = fc_conv with a CUDA event recorded around
 * every launch; unlike the other calls it synchronises the stream before returning. It writes the duration of
 * each launch (ms) to ms_out[0..*n_out), *n_out = min(max_n, n_launches). */
int fc_conv_profiled(const fc_plan* plan, const void* d_const, const float* d_x, const float* d_kspec, const float* d_bias,
                     float* d_y, void* d_ws, void* stream, float* ms_out, int max_n, int* n_out);
/* Name and algorithmic (compulsory read + write) bytes of launch i of fc_conv. */
int fc_plan_launch_info(const fc_plan* plan, int i, char* name, size_t namelen, int64_t* algo_bytes);

/* Standalone grouped per-bin contraction = reference complex_matmul(a, b, groups) (functional.py:11-16):
 * a: (B, Cin, bins) complex, b: (Cout, Cin/groups, bins) complex -> y: (B, Cout, bins) complex. No conjugation. */
int fc_complex_matmul(const float* d_a, const float* d_b, float* d_y, int64_t batch, int64_t cin, int64_t cout, int64_t groups,
                      int64_t bins, void* stream);

/* Tensor-core variant of the contraction for wide channel counts (tcgen05 kind::tf32 with a 3xTF32 split):
 * fc_tc_supported says whether the shape qualifies (Cin/groups >= 32 and a multiple of 16, Cout/groups a multiple of
 * 64, batch <= 80; fc_conv runs larger batches in chunks). The kernel spectrum is converted once by fc_tc_prepare_kernel into the bin-outermost planar
 * layout (same byte size as the input); d_scratch needs fc_tc_scratch_bytes. */
int fc_tc_supported(int64_t batch, int64_t cin, int64_t cout, int64_t groups);
int64_t fc_tc_scratch_bytes(int64_t batch, int64_t cin, int64_t cout, int64_t groups, int64_t bins);
int fc_tc_prepare_kernel(const float* d_kspec, float* d_kspec_tc, int64_t cin, int64_t cout, int64_t groups, int64_t bins, void* stream);
int fc_tc_complex_matmul(const float* d_a, const float* d_b_tc, float* d_y, void* d_scratch, int64_t batch, int64_t cin, int64_t cout,
                         int64_t groups, int64_t bins, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FFTCONV_B200_H */
