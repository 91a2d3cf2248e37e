/*
 * hgi_oracle.c -- CPU restatement of RustyHGI's hierarchical-grid encode/decode loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rustyhgi_b200/, include/) links,
 * loads or calls this file.  It is used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the *checker* and the CPU baseline.
 *
 * Parity pinning: the reference (Rust, nightly, un-vendored deps) cannot be compiled here
 * (no cargo/rustc), and its own unit tests hold no value-pinning vectors (src/lib.rs:61
 * shadows the source image, so :71-75 compares the decoded image with itself).  The oracle is
 * pinned against (a) the one golden artefact in the reference tree, docs/static_files/
 * lena_source.png -> lena_hgi.png (README.md:8), which this code reproduces bit-exactly with
 * `legacy_round=1` (tests/test_oracle_golden.py), and (b) an independent pure-Python
 * restatement (oracle/pyref.py).  HEAD's predictor differs from the artefact's only in the
 * final rounding term (src/interpolator.rs:51), which is restated literally below.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#define HGI_INTERP_CROSSED 0   /* src/interpolator.rs:6  (bincode variant index) */
#define HGI_INTERP_LEFTTOP 3   /* src/interpolator.rs:15 (no serialisation tag)  */
#define HGI_QUANT_NOOP   0     /* src/quantizator.rs:17 */
#define HGI_QUANT_LINEAR 1     /* src/quantizator.rs:36 */

typedef struct {
    uint8_t *buf;
    uint32_t width, height;
} plane_t;

/* src/quantizator.rs:41-63  Linear::from(QuantizationLevel) ; :19-23 NoOp::from */
void hgi_oracle_quant_table(int qkind, int qlevel, uint8_t table[256], uint8_t *error_out)
{
    static const uint8_t errors[4] = {0, 10, 20, 30};       /* quantizator.rs:43-48 */
    uint8_t error = (qkind == HGI_QUANT_LINEAR) ? errors[qlevel & 3] : 0;
    size_t scale = 2 * (size_t)error + 1;                   /* quantizator.rs:50 */
    for (size_t i = 0; i < 256; ++i) {
        if (qkind == HGI_QUANT_LINEAR) {
            size_t r = (i + error) / scale;                 /* quantizator.rs:52 */
            size_t v = r * scale;                           /* quantizator.rs:53 */
            table[i] = (uint8_t)v;                          /* quantizator.rs:54 `as u8` */
        } else {
            table[i] = (uint8_t)i;                          /* quantizator.rs:27-29 */
        }
    }
    if (error_out) *error_out = error;                      /* quantizator.rs:31-33,71-73 */
}

/* src/interpolator.rs:75-82  get_pixel closure: out-of-image reads as 0 */
static inline size_t get_pixel(const plane_t *im, uint32_t x, uint32_t y)
{
    if (x < im->width && y < im->height) return im->buf[(size_t)y * im->width + x];
    return 0;
}

/* src/interpolator.rs:57-91 (Crossed) and :15-28 (LeftTop).  `level` is the caller's level+1. */
static inline uint8_t interpolate(int interp, int legacy_round, uint32_t levels, uint32_t level,
                                  uint32_t x, uint32_t y, const plane_t *im)
{
    uint32_t step = 1u << (levels - level + 1);             /* interpolator.rs:19,67 */
    uint32_t mask = step - 1;
    uint32_t x_top = x - (x & mask);                        /* interpolator.rs:22,70 */
    uint32_t y_left = y - (y & mask);                       /* interpolator.rs:23,71 */
    if (interp == HGI_INTERP_LEFTTOP)
        return im->buf[(size_t)y_left * im->width + x_top]; /* interpolator.rs:26 */
    uint32_t x_bot = x_top + step;                          /* interpolator.rs:72 */
    uint32_t y_right = y_left + step;                       /* interpolator.rs:73 */
    size_t left_top  = get_pixel(im, x_top, y_left);        /* interpolator.rs:85 */
    size_t right_top = get_pixel(im, x_top, y_right);       /* interpolator.rs:86 */
    size_t left_bot  = get_pixel(im, x_bot, y_left);        /* interpolator.rs:87 */
    size_t right_bot = get_pixel(im, x_bot, y_right);       /* interpolator.rs:88 */
    /* interpolator.rs:43-54  CrossedValues::prediction */
    size_t left  = (left_top  + left_bot  + 1) >> 1;
    size_t right = (right_bot + right_top + 1) >> 1;
    size_t top   = (right_top + left_top  + 1) >> 1;
    size_t bot   = (right_bot + left_bot  + 1) >> 1;
    size_t sum = left + right + top + bot;
    /* HEAD: `>> 2` (interpolator.rs:51).  legacy_round is a TEST-ONLY switch that restores the
       `(sum+1)>>2` of the revision that produced docs/static_files/lena_hgi.png. */
    size_t average = legacy_round ? ((sum + 1) >> 2) : (sum >> 2);
    return (uint8_t)average;
}

/* src/encoder.rs:39-71  Encoder::encode.  `recon` is the by-value `input` the reference
   mutates; it holds the decoder-visible reconstruction on return. */
int hgi_oracle_encode_ex(const uint8_t *image, uint32_t width, uint32_t height, uint32_t levels,
                         int interp, int qkind, int qlevel, int legacy_round,
                         uint8_t *grid_out, uint8_t *recon_out, uint64_t *fixups_out)
{
    if (!image || !grid_out || levels > 30) return -1;
    size_t n = (size_t)width * height;
    uint8_t table[256];
    hgi_oracle_quant_table(qkind, qlevel, table, NULL);
    uint8_t *scratch = recon_out ? recon_out : (uint8_t *)malloc(n ? n : 1);
    if (!scratch) return -2;
    memcpy(scratch, image, n);                              /* `mut input: GrayImage` by value */
    plane_t input = {scratch, width, height};
    uint64_t fixups = 0;

    /* encoder.rs:26-37  initialize_first_level */
    uint64_t step0 = 1ull << levels;
    for (uint64_t line = 0; line < height; line += step0)
        for (uint64_t column = 0; column < width; column += step0)
            grid_out[line * width + column] = scratch[line * width + column];

    for (uint32_t level = 0; level < levels; ++level) {    /* encoder.rs:45 */
        /* src/utils.rs:11-41 traverse_level(level, levels, 0, width, 0, height, f) */
        uint32_t e = levels - level;                        /* utils.rs:16 */
        uint64_t step = 1ull << e;                          /* utils.rs:17 */
        uint64_t substep = 1ull << (e - 1);                 /* utils.rs:18 */
        uint64_t start = substep;                           /* utils.rs:19 (x1 = 0) */
        uint64_t line = 0;                                  /* utils.rs:21 */
#define PROCESS_PIXEL(column, line) do {                                                     \
            /* encoder.rs:46-65 */                                                           \
            uint8_t prediction = interpolate(interp, legacy_round, levels, level + 1,        \
                                             (uint32_t)(column), (uint32_t)(line), &input);  \
            size_t idx = (size_t)(line) * width + (size_t)(column);                          \
            uint8_t actual_value = scratch[idx];                     /* encoder.rs:52 */     \
            uint8_t diff = (uint8_t)(actual_value - prediction);     /* encoder.rs:53 */     \
            uint8_t quanted_diff = table[diff];                      /* encoder.rs:54 */     \
            int overflow = ((unsigned)prediction + quanted_diff) > 255;   /* :56 */          \
            int overflow_is_expected = ((unsigned)prediction + diff) > 255; /* :57 */        \
            if (overflow != overflow_is_expected) {                  /* encoder.rs:58 */     \
                quanted_diff = diff;                                 /* encoder.rs:59 */     \
                ++fixups;                                                                    \
            }                                                                                \
            grid_out[idx] = quanted_diff;                            /* encoder.rs:62 */     \
            scratch[idx] = (uint8_t)(prediction + quanted_diff);     /* encoder.rs:63-64 */  \
        } while (0)
        while (line < height) {                             /* utils.rs:22 */
            for (uint64_t column = start; column < width; column += step)   /* utils.rs:23-27 */
                PROCESS_PIXEL(column, line);
            line += substep;                                /* utils.rs:29 */
            if (line >= height) break;                      /* utils.rs:30-32 */
            for (uint64_t column = 0; column < width; column += substep)    /* utils.rs:34-38 */
                PROCESS_PIXEL(column, line);
            line += substep;                                /* utils.rs:39 */
        }
#undef PROCESS_PIXEL
    }
    if (fixups_out) *fixups_out = fixups;
    if (!recon_out) free(scratch);
    return 0;
}

int hgi_oracle_encode(const uint8_t *image, uint32_t width, uint32_t height, uint32_t levels,
                      int interp, int qkind, int qlevel, uint8_t *grid_out)
{
    return hgi_oracle_encode_ex(image, width, height, levels, interp, qkind, qlevel, 0,
                                grid_out, NULL, NULL);
}

/* src/decoder.rs:18-46  Decoder::decode */
int hgi_oracle_decode_ex(const uint8_t *grid, uint32_t width, uint32_t height, uint32_t levels,
                         int interp, int legacy_round, uint8_t *image_out)
{
    if (!grid || !image_out || levels > 30) return -1;
    size_t n = (size_t)width * height;
    memset(image_out, 0, n);                                /* decoder.rs:19 GrayImage::new */
    plane_t image = {image_out, width, height};

    uint64_t step0 = 1ull << levels;                        /* decoder.rs:22 */
    for (uint64_t line = 0; line < height; line += step0)   /* decoder.rs:23-28 */
        for (uint64_t column = 0; column < width; column += step0)
            image_out[line * width + column] = grid[line * width + column];

    for (uint32_t level = 0; level < levels; ++level) {    /* decoder.rs:30 */
        uint32_t e = levels - level;
        uint64_t step = 1ull << e;
        uint64_t substep = 1ull << (e - 1);
        uint64_t start = substep;
        uint64_t line = 0;
#define PROCESS_PIXEL(column, line) do {                                                     \
            /* decoder.rs:32-41 */                                                           \
            size_t idx = (size_t)(line) * width + (size_t)(column);                          \
            uint8_t diff = grid[idx];                                /* decoder.rs:33 */     \
            uint8_t prediction = interpolate(interp, legacy_round, levels, level + 1,        \
                                             (uint32_t)(column), (uint32_t)(line), &image);  \
            image_out[idx] = (uint8_t)(prediction + diff);           /* decoder.rs:39-40 */  \
        } while (0)
        while (line < height) {
            for (uint64_t column = start; column < width; column += step)
                PROCESS_PIXEL(column, line);
            line += substep;
            if (line >= height) break;
            for (uint64_t column = 0; column < width; column += substep)
                PROCESS_PIXEL(column, line);
            line += substep;
        }
#undef PROCESS_PIXEL
    }
    return 0;
}

int hgi_oracle_decode(const uint8_t *grid, uint32_t width, uint32_t height, uint32_t levels,
                      int interp, uint8_t *image_out)
{
    return hgi_oracle_decode_ex(grid, width, height, levels, interp, 0, image_out);
}

/* No reference code: exact 256-bin count of the residual-grid bytes. */
void hgi_oracle_histogram(const uint8_t *grid, size_t n, uint64_t hist_out[256])
{
    memset(hist_out, 0, 256 * sizeof(uint64_t));
    for (size_t i = 0; i < n; ++i) hist_out[grid[i]]++;
}

/* src/main.rs:84-92,106  `hgi test` squared-error accumulation; returns sd = sum / (w*h)
   (integer division, main.rs:106) and the raw sum / max abs error through pointers. */
uint64_t hgi_oracle_sd(const uint8_t *before, const uint8_t *after, size_t n,
                       uint64_t *sum_sq_out, uint32_t *max_abs_out)
{
    uint64_t sd = 0;
    uint32_t mx = 0;
    for (size_t i = 0; i < n; ++i) {
        int32_t d = (int32_t)before[i] - (int32_t)after[i]; /* main.rs:89 */
        uint32_t a = (uint32_t)(d < 0 ? -d : d);
        if (a > mx) mx = a;
        sd += (uint64_t)a * a;                              /* main.rs:91 */
    }
    if (sum_sq_out) *sum_sq_out = sd;
    if (max_abs_out) *max_abs_out = mx;
    return n ? sd / n : 0;                                  /* main.rs:106 */
}

/* The reference's f32 luma (image 0.19 `to_luma`, called at src/main.rs:42,74): third-party
   arithmetic, restated: l = 0.2126*r + 0.7152*g + 0.0722*b in f32, left to right, truncated. */
void hgi_oracle_rgb_to_luma(const uint8_t *rgb, size_t n_pixels, uint8_t *luma_out)
{
    for (size_t i = 0; i < n_pixels; ++i) {
        volatile float l = 0.2126f * (float)rgb[3 * i];
        l = l + 0.7152f * (float)rgb[3 * i + 1];
        l = l + 0.0722f * (float)rgb[3 * i + 2];
        luma_out[i] = (uint8_t)l;
    }
}

/* Batch drivers for the CPU baseline: the reference is single-threaded (one image per call);
   a by-image parallel loop (pthreads, dynamic work queue) is the most favourable way to run
   it on all host cores. */
#include <pthread.h>
#include <unistd.h>

typedef struct {
    const uint8_t *in;
    uint8_t *out;
    uint32_t n_images, width, height, levels;
    int interp, qkind, qlevel, decode;
    volatile int64_t next;
    volatile int rc;
} batch_job_t;

static void *batch_worker(void *arg)
{
    batch_job_t *job = (batch_job_t *)arg;
    size_t n = (size_t)job->width * job->height;
    for (;;) {
        int64_t i = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
        if (i >= (int64_t)job->n_images) break;
        int r = job->decode
            ? hgi_oracle_decode(job->in + (size_t)i * n, job->width, job->height, job->levels,
                                job->interp, job->out + (size_t)i * n)
            : hgi_oracle_encode(job->in + (size_t)i * n, job->width, job->height, job->levels,
                                job->interp, job->qkind, job->qlevel, job->out + (size_t)i * n);
        if (r) job->rc = r;
    }
    return NULL;
}

int hgi_oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static int run_batch(batch_job_t *job, int n_threads)
{
    if (n_threads <= 0) n_threads = hgi_oracle_max_threads();
    if (n_threads > 1024) n_threads = 1024;
    if ((uint32_t)n_threads > job->n_images) n_threads = (int)job->n_images;
    if (n_threads <= 1) { batch_worker(job); return job->rc; }
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    if (!tids) return -2;
    int started = 0;
    for (int t = 0; t < n_threads; ++t)
        if (pthread_create(&tids[started], NULL, batch_worker, job) == 0) ++started;
    if (started == 0) batch_worker(job);
    for (int t = 0; t < started; ++t) pthread_join(tids[t], NULL);
    free(tids);
    return job->rc;
}

int hgi_oracle_encode_batch(const uint8_t *images, uint32_t n_images, uint32_t width,
                            uint32_t height, uint32_t levels, int interp, int qkind, int qlevel,
                            uint8_t *grids_out, int n_threads)
{
    batch_job_t job = {images, grids_out, n_images, width, height, levels,
                       interp, qkind, qlevel, 0, 0, 0};
    return run_batch(&job, n_threads);
}

int hgi_oracle_decode_batch(const uint8_t *grids, uint32_t n_images, uint32_t width,
                            uint32_t height, uint32_t levels, int interp, uint8_t *images_out,
                            int n_threads)
{
    batch_job_t job = {grids, images_out, n_images, width, height, levels,
                       interp, 0, 0, 1, 0, 0};
    return run_batch(&job, n_threads);
}
