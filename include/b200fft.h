/* b200fft -- C ABI of the B200-native batched complex FFT backend.
 *
 * This is the drop-in boundary for RustFFT's FftPlanner / Fft::process hot path: the entry points
 * a `src/cuda/` backend of the reference (next to src/avx, src/sse) would bind over FFI.  Plain
 * pointers and sizes only.  Citations are relative to the reference tree (RustFFT 6.4.1).
 *
 *   reference interface                                     replaced by
 *   ------------------------------------------------------  ---------------------------------------
 *   FftPlannerAvx::new() -> Result<Self,()>                  b200fft_device_count()
 *     (src/avx/avx_planner.rs:121-164, probed in             (0 devices => the Rust shim's new() returns
 *      FftPlanner::new, src/plan.rs:72-94)                    Err(()) and the next backend is tried)
 *   FftPlanner::plan_fft(len, direction) -> Arc<dyn Fft<T>>  b200fft_plan_create() / b200fft_plan_destroy()
 *   FftPlannerScalar::design_fft_for_len -> Recipe           b200fft_plan_create_from_recipe()  (the host keeps
 *     (src/plan.rs:134-226,412-425) + build_fft (:315-410)     planning; the library builds what the recipe says)
 *     (src/plan.rs:101-126; instances cached per             (the shim keeps RustFFT's FftCache,
 *      (len, direction), src/fft_cache.rs:5-38)               src/fft_cache.rs, above this call)
 *   Length::len / Direction::fft_direction                   b200fft_plan_len() / b200fft_plan_direction()
 *     (src/lib.rs:140-181)
 *   Fft::get_{inplace,outofplace,immutable}_scratch_len      b200fft_plan_scratch_len()   (always 0: a backend
 *     (src/lib.rs:262-277)                                    may report 0, src/lib.rs:259-261)
 *   Fft::process / process_with_scratch (in place)           b200fft_exec_host_inplace()
 *     (src/lib.rs:195-211)
 *   Fft::process_outofplace_with_scratch /                   b200fft_exec_host_outofplace()
 *   Fft::process_immutable_with_scratch (src/lib.rs:231-255)
 *   -- (no reference equivalent: device-resident batch)      b200fft_exec_device() / b200fft_exec_device_ws()
 *
 * Semantics kept from the reference:
 *   - unnormalised, natural order, forward sign exp(-2*pi*i*k*n/N)   (src/lib.rs:81-89, src/twiddles.rs:11)
 *   - Complex<T> is repr(C) {re, im}: a buffer is float2[] / double2[], interleaved       (CHANGELOG.md:139)
 *   - a buffer of n_complex = batch*len elements is `batch` independent contiguous transforms
 *                                                                     (src/array_utils.rs:151-177)
 *   - len == 0 is a silent no-op (src/fft_helper.rs:16-18); len 0 and 1 plan fine (src/plan.rs:873-882)
 *   - the reference PANICS on bad arguments (src/common.rs:13-104); here every call returns a status and
 *     b200fft_last_error() returns the same message text; the Rust shim turns non-zero into panic!().
 *   - plan handles are immutable and may be used concurrently from many threads, like Arc<dyn Fft<T>>
 *     (src/lib.rs:184 `Sync + Send`, examples/concurrency.rs:17-29).
 *
 * There is NO CPU fallback: without a CUDA device every plan/exec call fails with B200FFT_ERR_NO_DEVICE.
 */
#ifndef B200FFT_H
#define B200FFT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200fft_plan b200fft_plan;

enum {
    B200FFT_OK = 0,
    B200FFT_ERR_INVALID_ARG = -1,   /* null pointer, unknown enum value */
    B200FFT_ERR_NO_DEVICE = -2,     /* no CUDA device / not an sm_100 part */
    B200FFT_ERR_CUDA = -3,          /* a CUDA runtime call failed; see b200fft_last_error() */
    B200FFT_ERR_BUFFER_TOO_SMALL = -4, /* "Provided FFT buffer was too small" (src/common.rs:19-24) */
    B200FFT_ERR_NOT_MULTIPLE = -5,  /* "Input FFT buffer must be a multiple of FFT length" (src/common.rs:25-31) */
    B200FFT_ERR_LEN_MISMATCH = -6,  /* "input buffer and output buffer must have the same length" (src/common.rs:50) */
    B200FFT_ERR_UNSUPPORTED = -7,   /* length outside what this build can plan */
    B200FFT_ERR_WORKSPACE = -8      /* caller-provided workspace too small */
};

enum { B200FFT_FORWARD = 0, B200FFT_INVERSE = 1 };  /* FftDirection, src/lib.rs:147-171 */
enum { B200FFT_F32 = 0, B200FFT_F64 = 1 };          /* Complex<f32> / Complex<f64> */

/* Number of usable (compute capability 10.x) devices.  Returns status; *n = 0 when there is none. */
int b200fft_device_count(int* n);

/* Build a plan: picks the kernel sequence for `len`, builds twiddle / chirp / index tables on the host
 * (angles evaluated in extended precision, rounded once -- src/twiddles.rs:6-23 contract) and uploads them. */
int b200fft_plan_create(b200fft_plan** out, uint64_t len, int direction, int precision, int device);
int b200fft_plan_destroy(b200fft_plan* plan);

/* Planning owned by the caller (the Rust host code of a `src/cuda/` backend keeps src/plan.rs's design_fft_* and hands the
 * decomposition over as data): a flattened recipe tree, node 0 = the root, in the vocabulary of the reference's Recipe enum
 * (src/plan.rs:134-226).  The library maps each node onto its kernels or returns B200FFT_ERR_UNSUPPORTED; tables are always
 * recomputed here in extended precision (src/twiddles.rs:6-23 contract), so no host twiddles cross the boundary.
 *   kind          reference Recipe                      a, b, child
 *   AUTO          --                                    this library's own choice for `len`
 *   POW2          Radix4 / Butterfly2..32               --  (len = 2^k: one CTA pass up to 2^14, two passes up to 2^24)
 *   SMOOTH        RadixN / Butterfly3..31               --  (prime factors <= 31, one CTA pass)
 *   MIXED_RADIX   MixedRadix / MixedRadixSmall          len = a * b, two passes (a x b)
 *   GOOD_THOMAS   GoodThomasAlgorithm(+Small)           len = a * b, gcd(a, b) = 1, two passes without twiddles
 *   RADER         RadersAlgorithm                       len = a * p (a = 0 or 1: len = p prime; a in 2..8: MixedRadix{a x Rader(p)}
 *                                                       fused); child = node of the inner FFT of length p - 1 (0 = AUTO)
 *   BLUESTEIN     BluesteinsAlgorithm                   child = node of the inner FFT, length M >= 2 len - 1 (0 = AUTO)
 *   CLUSTER       MixedRadix, on chip                   len = 2^14 .. 2^17 (f32): both passes inside one thread-block cluster, the
 *                                                       transpose through distributed shared memory (one pass over HBM);
 *                                                       a = 1: half tiles (4096 points per CTA, 2^14 .. 2^16) */
enum {
    B200FFT_RECIPE_AUTO = 0,
    B200FFT_RECIPE_POW2 = 1,
    B200FFT_RECIPE_SMOOTH = 2,
    B200FFT_RECIPE_MIXED_RADIX = 3,
    B200FFT_RECIPE_GOOD_THOMAS = 4,
    B200FFT_RECIPE_RADER = 5,
    B200FFT_RECIPE_BLUESTEIN = 6,
    B200FFT_RECIPE_CLUSTER = 7,
    B200FFT_RECIPE_COLUMNS = 8 /* internal to the 2-D plans: len = a * b, a-point FFTs down the columns of [a][b] images */
};
typedef struct b200fft_recipe_node {
    uint32_t kind;  /* B200FFT_RECIPE_* */
    uint32_t child; /* index of the inner-FFT node (RADER / BLUESTEIN); 0 = let the library choose */
    uint64_t len;   /* length of this node's transform */
    uint64_t a, b;  /* MIXED_RADIX / GOOD_THOMAS: the split; RADER: a = outer radix */
} b200fft_recipe_node;
int b200fft_plan_create_from_recipe(b200fft_plan** out, const b200fft_recipe_node* nodes, uint32_t n_nodes, int direction,
                                    int precision, int device);

/* The decomposition a plan was built as, in the same node vocabulary (node 0 = root): a plan can be stored as data and rebuilt
 * bit-identically with b200fft_plan_create_from_recipe -- plan serialisation for callers that cache plans across processes
 * (the reference keeps its Recipe in memory only, src/plan.rs:134-226).  Returns the number of nodes (at most 2 today); writes
 * min(cap, n) of them when `nodes` is not NULL. */
int b200fft_plan_recipe(const b200fft_plan* plan, b200fft_recipe_node* nodes, uint32_t cap);

uint64_t b200fft_plan_len(const b200fft_plan* plan);
int b200fft_plan_direction(const b200fft_plan* plan);
int b200fft_plan_precision(const b200fft_plan* plan);
/* which: 0 = in place, 1 = out of place, 2 = immutable.  Host-visible scratch the caller must supply: 0. */
uint64_t b200fft_plan_scratch_len(const b200fft_plan* plan, int which);
/* Human-readable description of the chosen algorithm tree, e.g. "FourStep{256x256}".  Returns length or <0. */
int b200fft_plan_describe(const b200fft_plan* plan, char* buf, uint64_t cap);
/* Kernel launches one exec of `batch` transforms issues (bench.py reports it as gpu_launches). */
uint64_t b200fft_plan_launches(const b200fft_plan* plan, uint64_t batch);

/* Trait-conformant host paths: host memory owned by the caller (pageable or pinned), n_complex = batch*len elements;
 * synchronous.  The batch flows in 64 MiB slices through a three-stage pipeline (copy in | kernels | copy out, one stream each)
 * over a ring of four device buffers owned by the plan (created on the first call, reused afterwards).  Pinned / registered
 * memory is copied from and to directly; pageable memory is staged through a pinned ring by a small pool of copy threads. */
int b200fft_exec_host_inplace(const b200fft_plan* plan, void* buffer, uint64_t n_complex);
int b200fft_exec_host_outofplace(const b200fft_plan* plan, const void* input, void* output, uint64_t n_complex);

/* Device-resident path (the measured one): d_in / d_out hold batch*len elements on the plan's device,
 * d_in == d_out allowed; asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream).
 * Plans that need an intermediate buffer take it from the stream-ordered allocator. */
int b200fft_exec_device(const b200fft_plan* plan, const void* d_in, void* d_out, uint64_t batch, void* cuda_stream);
/* Same, with a caller-provided device workspace of at least b200fft_workspace_bytes(plan, batch) bytes, 128-byte aligned
 * (the passes address it through TMA and drop consumed lines with discard.global.L2; B200FFT_ERR_INVALID_ARG otherwise). */
uint64_t b200fft_workspace_bytes(const b200fft_plan* plan, uint64_t batch);
int b200fft_exec_device_ws(const b200fft_plan* plan, const void* d_in, void* d_out, uint64_t batch,
                           void* cuda_stream, void* d_workspace, uint64_t workspace_bytes);

/* Real-input / real-output transforms of even length on top of the complex plans (what the `realfft` crate adds above
 * RustFFT's Fft trait; RustFFT itself has none).  forward: batch * len reals -> batch * (len/2 + 1) complex (the non-redundant half of the
 * spectrum); inverse: the reverse, unnormalised (inverse(forward(x)) = len * x).  Device-resident entry points (asynchronous on the stream)
 * and synchronous host ones (plain copies in and out, not pipelined). */
typedef struct b200fft_real_plan b200fft_real_plan;
int b200fft_real_plan_create(b200fft_real_plan** out, uint64_t len, int precision, int device);
int b200fft_real_plan_destroy(b200fft_real_plan* plan);
uint64_t b200fft_real_workspace_bytes(const b200fft_real_plan* plan, uint64_t batch);
int b200fft_real_forward_device(const b200fft_real_plan* plan, const void* d_real_in, void* d_complex_out, uint64_t batch, void* cuda_stream);
int b200fft_real_inverse_device(const b200fft_real_plan* plan, const void* d_complex_in, void* d_real_out, uint64_t batch, void* cuda_stream);
int b200fft_real_forward_host(const b200fft_real_plan* plan, const void* real_in, void* complex_out, uint64_t batch);
int b200fft_real_inverse_host(const b200fft_real_plan* plan, const void* complex_in, void* real_out, uint64_t batch);

/* 2-D complex transforms of row-major [height][width] images (batch of them, contiguous): the width-point plan over all rows, then one
 * strided pass of height-point FFTs down the columns (the same two steps a caller of RustFFT writes with two plans and two
 * transposes).  height: prime factors <= 31 and at most 4096 (f64: 2048); width: any length b200fft_plan_create accepts. */
typedef struct b200fft_plan2d b200fft_plan2d;
int b200fft_plan2d_create(b200fft_plan2d** out, uint64_t height, uint64_t width, int direction, int precision, int device);
int b200fft_plan2d_destroy(b200fft_plan2d* plan);
int b200fft_exec2d_device(const b200fft_plan2d* plan, const void* d_in, void* d_out, uint64_t batch, void* cuda_stream);
int b200fft_exec2d_host(const b200fft_plan2d* plan, const void* in, void* out, uint64_t batch);

/* Message of the last failing call on this thread ("" if none). */
const char* b200fft_last_error(void);
/* Library build string: "b200fft <version> sm_100a" */
const char* b200fft_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200FFT_H */
