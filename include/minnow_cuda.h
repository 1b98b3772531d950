/*
 * minnow_cuda.h -- C ABI of libminnow_b200.so: the block encode/decode hot path
 * of phil-mansfield/minnow on one NVIDIA B200 (sm_100a).
 *
 * The reference has no FFI today: its seam for this path is the private Go
 * `group` interface (go/group.go:77-87) and package `bit`
 * (go/bit/bit.go:19-206), called from Writer.Data (go/writer.go:90-104),
 * Writer.Close (:107-141), Reader.Open (go/reader.go:28-88) and Reader.Data
 * (:114-127).  Each entry point below names the reference code it replaces;
 * INTEGRATION.md shows the cgo stubs that bind them behind the unchanged Go
 * API.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * Conventions
 *  - every function returns 0 (MNW_OK) or a negative mnw_status; the message is
 *    at mnw_last_error(ctx) (the Go glue re-panics with it, because the
 *    reference's error convention is panic: go/writer.go:34,92; go/bit/bit.go:31).
 *  - a context owns one CUDA stream, pinned staging and device scratch that grow
 *    like bit.ArrayBuffer (go/bit/bit.go:188-206).  Like a minnow.Writer, a
 *    context is NOT safe for concurrent use; use one per goroutine/thread.
 *  - caller owns all in/out buffers for the duration of the call only.
 *  - `_dev` variants take DEVICE pointers, enqueue on the context's stream and
 *    do not synchronise; call mnw_sync() before reading results.
 *  - integers are little-endian int64 like every on-disk integer of the format.
 *  - there is no CPU fallback: without a CUDA device mnw_create fails.
 *
 * Results are bit-identical to the Go reference on the same inputs: packed
 * bytes, per-block (min, bits), byte offsets, decoded int64 values, decoded
 * pixel indices; decoded float32 values are bit-identical GIVEN the jitter
 * stream (see mnw_jitter below).
 */
#ifndef MINNOW_CUDA_H
#define MINNOW_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNW_API __attribute__((visibility("default")))

typedef enum {
    MNW_OK = 0,
    MNW_ERR_CUDA = -1,       /* CUDA runtime error (message has the detail)  */
    MNW_ERR_ARG = -2,        /* bad argument (where the reference panics)     */
    MNW_ERR_CAPACITY = -3,   /* output buffer too small                       */
    MNW_ERR_FORMAT = -4,     /* not a minnow/minh/minp file, wrong version    */
    MNW_ERR_IO = -5,         /* file I/O failed                               */
    MNW_ERR_TYPE = -6        /* TypeMatch failure, go/group.go:43-71          */
} mnw_status;

/* Group type codes, identical to go/group.go:11-24. */
enum {
    MNW_INT64_GROUP = 0, MNW_INT32_GROUP, MNW_INT16_GROUP, MNW_INT8_GROUP,
    MNW_UINT64_GROUP, MNW_UINT32_GROUP, MNW_UINT16_GROUP, MNW_UINT8_GROUP,
    MNW_FLOAT64_GROUP, MNW_FLOAT32_GROUP, MNW_INT_GROUP, MNW_FLOAT_GROUP
};

typedef struct mnw_ctx mnw_ctx;

/* ---- context ------------------------------------------------------------ */
MNW_API int mnw_create(int device, mnw_ctx **out);
MNW_API void mnw_destroy(mnw_ctx *ctx);
MNW_API const char *mnw_last_error(const mnw_ctx *ctx); /* ctx may be NULL: last create error */
MNW_API int mnw_sync(mnw_ctx *ctx);                     /* wait for the context's stream */
MNW_API void *mnw_stream(mnw_ctx *ctx);                 /* the cudaStream_t, for event timing */
MNW_API const char *mnw_version(void);
/* Number of kernels this library has launched on ctx since creation. */
MNW_API int64_t mnw_launch_count(const mnw_ctx *ctx);

/* ---- package bit (leaf parity API) ----------------------------------------
 * mnw_precision_needed  = bit.PrecisionNeeded  go/bit/bit.go:19-21 (Go's
 *                         float64 log2 rounding is reproduced, including its
 *                         under-count from max = 2^49; max = 2^64-1 returns
 *                         MNW_ERR_ARG where Go's result is undefined).
 * mnw_array_bytes       = bit.ArrayBytes       go/bit/bit.go:23-25
 * mnw_pack              = bit.BufferedArray    go/bit/bit.go:84-134 (HOST ptrs;
 *                         out holds mnw_array_bytes(bits, n); bits in 1..64)
 * mnw_unpack            = (*Array).Slice       go/bit/bit.go:29-82  (HOST ptrs)
 * mnw_bits              = ArrayBuffer.Bits     go/bit/bit.go:151-159 (HOST ptr)
 */
MNW_API int mnw_precision_needed(uint64_t max);
MNW_API int64_t mnw_array_bytes(int bits, int64_t n);
MNW_API int mnw_pack(mnw_ctx *ctx, int bits, const uint64_t *x, int64_t n, uint8_t *out);
MNW_API int mnw_unpack(mnw_ctx *ctx, int bits, const uint8_t *in, int64_t n, uint64_t *out);
MNW_API int mnw_bits(mnw_ctx *ctx, const uint64_t *x, int64_t n, int *bits);

/* ---- group codecs ---------------------------------------------------------
 * One call encodes nblocks consecutive blocks of one group; this replaces
 * nblocks calls of intGroup.writeData (go/group.go:242-255) or
 * floatGroup.writeData (:312-327) and the addBlock prefix sum
 * (go/block_index.go:16-23).  Outputs, all [nblocks]:
 *   mins[b], bits[b]  = what the reference appends to g.mins / g.bits
 *   offsets[b]        = blockOffset(startBlock + b), go/block_index.go:25-35
 *                       (EXCLUSIVE prefix sum of ArrayBytes(bits[b], n_b))
 *   out[0, *out_len)  = the bytes the reference writes to the file for the
 *                       group, blocks back to back (a block of 0 bits is empty)
 * Blocks may be ragged: block b holds elements [starts[b], starts[b+1]) of x;
 * pass starts = NULL for the format's usual equal-size case (block b =
 * [b*n, (b+1)*n)); with starts given, n is ignored.
 */
typedef struct {
    float low, high;      /* floatGroup.low / .high, go/group.go:271             */
    int64_t pixels;       /* floatGroup.pixels (mnw_float_group_pixels)          */
    uint8_t periodic;     /* floatGroup.periodic; Writer.FloatGroup passes 1     */
    uint8_t log10;        /* minh Column.Log != 0: x <- float32(log10(float64 x)) */
    uint8_t clamp;        /* minh processFloatGroup clamp to [low, nextafter(high)) */
    uint8_t reserved[5];
} mnw_float_desc;

/* go/writer.go:73: pixels = int64(ceil(float64((hi - lo) / dx))) in float32 */
MNW_API int64_t mnw_float_group_pixels(float lo, float hi, float dx);

MNW_API int mnw_encode_int_group(mnw_ctx *ctx, const int64_t *x, int64_t n, int64_t nblocks,
                                 const int64_t *starts, int64_t *mins, int64_t *bits,
                                 int64_t *offsets, uint8_t *out, int64_t out_cap, int64_t *out_len);
MNW_API int mnw_encode_float_group(mnw_ctx *ctx, const mnw_float_desc *desc, const float *x,
                                   int64_t n, int64_t nblocks, const int64_t *starts,
                                   int64_t *mins, int64_t *bits, int64_t *offsets,
                                   uint8_t *out, int64_t out_cap, int64_t *out_len);

/* Gathered blocks: block b holds col[idx[i]] for i in [starts[b], starts[b+1]) -- the per-cell
 * column gather of BoundaryWriter.Column / boundaryColumn (go/minh/boundary.go:184-256) fused with
 * the group encode.  col has ncol elements; every idx must lie in [0, ncol). */
MNW_API int mnw_encode_int_group_gather(mnw_ctx *ctx, const int64_t *col, int64_t ncol, const int64_t *idx,
                                        int64_t nblocks, const int64_t *starts, int64_t *mins, int64_t *bits,
                                        int64_t *offsets, uint8_t *out, int64_t out_cap, int64_t *out_len);
MNW_API int mnw_encode_float_group_gather(mnw_ctx *ctx, const mnw_float_desc *desc, const float *col, int64_t ncol,
                                          const int64_t *idx, int64_t nblocks, const int64_t *starts,
                                          int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out,
                                          int64_t out_cap, int64_t *out_len);

/* All IntGroup / FloatGroup columns of ONE minh block in one call: the per-column loop of
 * minh.Writer.Block (go/minh/minh.go:99-139, processFloatGroup :141-149 through desc.log10 / desc.clamp).
 * Column c holds n values at data[c] (int64 when is_float == 0, float32 otherwise) and becomes one block of
 * its own minnow group: mins[c], bits[c], nbytes[c] = what intGroup / floatGroup.writeData record, packed
 * bytes at out + c * out_col_stride.  FloatGroup columns must be periodic (Writer.FloatGroup always is). */
typedef struct {
    int32_t is_float;
    int32_t reserved;
    mnw_float_desc desc;   /* FloatGroup columns only */
} mnw_column;
MNW_API int mnw_encode_columns(mnw_ctx *ctx, int64_t ncols, const mnw_column *cols, const void *const *data,
                               int64_t n, int64_t *mins, int64_t *bits, int64_t *nbytes, uint8_t *out,
                               int64_t out_col_stride);

/* The same on device-resident columns: data_dev is a HOST array of ncols DEVICE pointers; mins, bits, nbytes, out are
 * DEVICE pointers; enqueued on the context's stream without synchronising (errors: mnw_sync). */
MNW_API int mnw_encode_columns_dev(mnw_ctx *ctx, int64_t ncols, const mnw_column *cols, const void *const *data_dev,
                                   int64_t n, int64_t *mins, int64_t *bits, int64_t *nbytes, uint8_t *out,
                                   int64_t out_col_stride);

/* ---- text -> typed columns: the producer of minh.Writer.Block's input (scripts/text_to_minh.go:166-214) ------------------
 * mnw_text_parse_block = the body of text.Reader.Block (go/text/text.go:181-200) for one block of bytes (HOST buf): split at
 * '\n', uncomment, trim, fields (go/text/parse.go:16-79,175-211), then strconv.Atoi of the n_icols columns icols[] and
 * float32(strconv.ParseFloat(.., 64)) of the n_fcols columns fcols[] (ascending column numbers, at most 64 each; :81-172).
 * The parsed columns STAY ON THE DEVICE as row-major matrices [n_icols][nrows] int64 and [n_fcols][nrows] float32 --
 * mnw_text_columns_dev hands out their device addresses (for mnw_encode_columns_dev: text to minh without the columns
 * ever visiting the host), mnw_text_columns copies them to HOST arrays.  Decimal -> float64 is correctly rounded (exact
 * fast path, else Eisel-Lemire with Go's own power-of-ten table); the rare fields that path cannot decide are returned
 * in `fallback` as triples (row, index into fcols, byte offset | length << 40) with 0 written in their place, for the
 * caller to convert with the host language's parser (*nfallback of them).  Where the reference panics (a field that is
 * not a number, a line with another column count, a requested column beyond the data) the call fails. */
MNW_API int mnw_text_parse_block(mnw_ctx *ctx, const char *buf, int64_t len, char sep, char comment, int n_icols,
                                 const int *icols, int n_fcols, const int *fcols, int64_t *nrows, int64_t *nfallback);
MNW_API int mnw_text_columns(mnw_ctx *ctx, int64_t *iout, float *fout, int64_t *fallback);
MNW_API int mnw_text_columns_dev(mnw_ctx *ctx, const int64_t **icols_dev, const float **fcols_dev);

/* ---- Lagrangian re-gridding: the producer of minp.Writer.Vectors' input ------------------------------------------------
 * vectorGrid.Insert over a batch (go/minp/snapshot/grid.go:206-211 with grid.Index :118-137, driven by xGrid / vGrid
 * :168-204): particle j with 1-based ID ids[j] goes to cell c, slot i of a Lagrangian lattice of ncell^3 cells x nside^3
 * particles.  grid_dev is a DEVICE array [ncell^3][nside^3][3] float32 that persists across calls (a snapshot arrives
 * file by file) and is exactly the input layout of mnw_minp_encode_vectors_dev (nfiles = ncell^3, nfile = nside).
 * mnw_regrid_insert takes HOST ids / vec and synchronises; the _dev form takes DEVICE pointers and only enqueues.
 * An ID outside [1, (ncell * nside)^3] is MNW_ERR_ARG (grid.Index panics). */
MNW_API int mnw_regrid_insert(mnw_ctx *ctx, const int64_t *ids, const float *vec, int64_t n, int64_t ncell, int64_t nside,
                              float *grid_dev);
MNW_API int mnw_regrid_insert_dev(mnw_ctx *ctx, const int64_t *ids, const float *vec, int64_t n, int64_t ncell,
                                  int64_t nside, float *grid_dev);

/* ---- minh BoundaryWriter (go/minh/boundary.go): cell + ghost-layer binning and the per-cell column encode -------------
 * mnw_boundary_coordinates = BoundaryWriter.Coordinates (:39-51): cellSizes (:93-109), indices (:54-86), hostCells
 *   (:111-151), idxReg (:154-165) and region (:173-180) for n points (HOST x, y, z) in a periodic box of size L cut into
 *   cells^3 cells with ghost layers of width `boundary`.  sizes [cells^3] (HOST) gets the number of points of every
 *   cell + ghost region, total their sum.  The per-cell index lists and boundary flags -- in the reference's order -- STAY
 *   ON THE DEVICE in the context for the column calls below; mnw_boundary_index copies them out (idx, flags: HOST
 *   int64 [total]; flags are 0 in a point's own cell and 1 in a ghost region).  A coordinate outside [0, 2 L) is
 *   MNW_ERR_ARG (the reference indexes outside its grid and panics).
 * mnw_boundary_encode_{int,float}_column = the IntGroup / FloatGroup branches of BoundaryWriter.Column (:184-225): cell i
 *   becomes block i (its own minnow group in the file) holding col[idx] of the cell's points, gathered on the device.
 * mnw_boundary_encode_flags = boundaryColumn (:227-246).  Outputs as in mnw_encode_int_group (cells^3 blocks). */
MNW_API int mnw_boundary_coordinates(mnw_ctx *ctx, const float *x, const float *y, const float *z, int64_t n, float L,
                                     float boundary, int64_t cells, int64_t *sizes, int64_t *total);
MNW_API int mnw_boundary_index(mnw_ctx *ctx, int64_t *idx, int64_t *flags);
MNW_API int mnw_boundary_encode_int_column(mnw_ctx *ctx, const int64_t *col, int64_t ncol, int64_t *mins, int64_t *bits,
                                           int64_t *offsets, uint8_t *out, int64_t out_cap, int64_t *out_len);
MNW_API int mnw_boundary_encode_float_column(mnw_ctx *ctx, const mnw_float_desc *desc, const float *col, int64_t ncol,
                                             int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out,
                                             int64_t out_cap, int64_t *out_len);
MNW_API int mnw_boundary_encode_flags(mnw_ctx *ctx, int64_t *mins, int64_t *bits, int64_t *offsets, uint8_t *out,
                                      int64_t out_cap, int64_t *out_len);

/* Decode nsel blocks of one group.  data/offsets/mins/bits describe the whole
 * group (nblocks entries; offsets as produced above); sel lists the block ids
 * to decode (NULL = blocks 0..nsel-1); block sel[j] lands at out + j*n.
 * Replaces intGroup.readData (go/group.go:257-263) / floatGroup.readData
 * (:299-310) per selected block (all blocks have n elements here).
 *
 * mnw_jitter: floatGroup.readData adds rand.Float64() from Go's global source
 * (go/group.go:308), so the reference's decoded floats are not reproducible.
 * The stream is explicit here:
 *   mode MNW_JITTER_CENTER: u = 0.5
 *   mode MNW_JITTER_HASH  : u = (mnw_jitter_hash32(seed, block_id, i) >> 8) * 2^-24,
 *                           block_id = block_id0 + sel[j]  (24 bits: u is exact in
 *                           float32 and float64(q) + u is exact for q < 2^29)
 *   mode MNW_JITTER_STREAM: u = u_stream[j*n + i]   (caller's doubles in [0,1))
 * out = dx*float32(float64(q) + u) + low in float32 without FMA, as in Go.
 */
enum { MNW_JITTER_CENTER = 0, MNW_JITTER_HASH = 1, MNW_JITTER_STREAM = 2 };
typedef struct {
    int32_t mode;
    int32_t reserved;
    uint64_t seed;
    uint64_t block_id0;
    const double *u_stream;
} mnw_jitter;
MNW_API uint32_t mnw_jitter_hash32(uint64_t seed, uint64_t block_id, uint64_t i);

MNW_API int mnw_decode_int_blocks(mnw_ctx *ctx, const uint8_t *data, int64_t data_len,
                                  const int64_t *offsets, const int64_t *mins, const int64_t *bits,
                                  int64_t n, int64_t nsel, const int64_t *sel, int64_t *out);
MNW_API int mnw_decode_float_blocks(mnw_ctx *ctx, const mnw_float_desc *desc, const uint8_t *data,
                                    int64_t data_len, const int64_t *offsets, const int64_t *mins,
                                    const int64_t *bits, int64_t n, int64_t nsel, const int64_t *sel,
                                    const mnw_jitter *jitter, float *out);

/* ---- minp: fused sub-cell gather + 3-axis encode --------------------------
 * Replaces the body of minp.Writer.Vectors (go/minp/minp.go:112-118): for each
 * axis k and sub-cell sc, getSubCell (:246-264) + floatGroup.writeData.  aos is
 * the [nfile^3][3]float32 Lagrangian cube of one file; desc[k] are the three
 * FloatGroups' parameters.  Block (k, sc) is entry k*subcells^3 + sc of
 * mins/bits/offsets (offsets restart at 0 for every axis = every group);
 * axis k's bytes are written to out + out_axis_stride*k and its length to
 * out_len[k].
 * mnw_decode_vec3_subcells replaces minp.Reader.Vectors (:191-206): decode,
 * periodic wrap (x<0 -> x+L; x>=L -> x-L when wrap_L > 0), setSubCell (:270-288).
 * jitter block ids are block_id0 + k*subcells^3 + sc.
 */
MNW_API int mnw_encode_vec3_subcells(mnw_ctx *ctx, const mnw_float_desc desc[3], const float *aos,
                                     int64_t nfile, int64_t subcells, int64_t *mins, int64_t *bits,
                                     int64_t *offsets, uint8_t *out, int64_t out_axis_stride,
                                     int64_t out_len[3]);
/* The whole of minp.Writer.Vectors for one file (go/minp/minp.go:86-119) with ONE upload of the
 * particles: limits ([0, L) per axis when periodic, go/minp/minp.go:88-90; else bounds() and
 * Nextafter32, :92-95), pixels (go/writer.go:73), then the encode above.  desc_out receives the
 * three FloatGroup parameter sets the file's group tails need. */
MNW_API int mnw_minp_encode_vectors(mnw_ctx *ctx, const float *aos, int64_t nfile, int64_t subcells, int periodic,
                                    float L, float dx, mnw_float_desc desc_out[3], int64_t *mins, int64_t *bits,
                                    int64_t *offsets, uint8_t *out, int64_t out_axis_stride, int64_t out_len[3]);
MNW_API int mnw_decode_vec3_subcells(mnw_ctx *ctx, const mnw_float_desc desc[3],
                                     const uint8_t *const data[3], const int64_t data_len[3],
                                     const int64_t *offsets, const int64_t *mins, const int64_t *bits,
                                     int64_t nfile, int64_t subcells, float wrap_L,
                                     const mnw_jitter *jitter, float *aos_out);

/* ---- block index ----------------------------------------------------------
 * Exclusive prefix sum of per-block byte sizes = blockIndex.addBlock +
 * blockOffset (go/block_index.go:16-35).  HOST pointers.  total gets the sum. */
MNW_API int mnw_scan_offsets(mnw_ctx *ctx, const int64_t *nbytes, int64_t nblocks, int64_t base,
                             int64_t *offsets, int64_t *total);

/* ---- device-resident variants (benchmarks, multi-GPU sharding) -------------
 * Same semantics; every pointer is a DEVICE pointer except desc/jitter structs
 * (host) and out_len (device int64).  Enqueued on the context's stream. */
MNW_API int mnw_encode_int_group_dev(mnw_ctx *ctx, const int64_t *x, int64_t n, int64_t nblocks,
                                     int64_t *mins, int64_t *bits, int64_t *offsets,
                                     uint8_t *out, int64_t out_cap, int64_t *out_len);
MNW_API int mnw_encode_float_group_dev(mnw_ctx *ctx, const mnw_float_desc *desc, const float *x,
                                       int64_t n, int64_t nblocks, int64_t *mins, int64_t *bits,
                                       int64_t *offsets, uint8_t *out, int64_t out_cap,
                                       int64_t *out_len);
MNW_API int mnw_decode_int_blocks_dev(mnw_ctx *ctx, const uint8_t *data, int64_t data_len,
                                      const int64_t *offsets, const int64_t *mins, const int64_t *bits,
                                      int64_t n, int64_t nsel, const int64_t *sel, int64_t *out);
MNW_API int mnw_decode_float_blocks_dev(mnw_ctx *ctx, const mnw_float_desc *desc, const uint8_t *data,
                                        int64_t data_len, const int64_t *offsets, const int64_t *mins,
                                        const int64_t *bits, int64_t n, int64_t nsel,
                                        const int64_t *sel, const mnw_jitter *jitter, float *out);

/* The read side of the same block: every IntGroup / FloatGroup column in at most three launches (plain float, Log, int: the per-column loop of
 * minh.Reader.Block, go/minh/minh.go:296-323; Log columns come back through 10^x).  Column c is block 0 of its own group,
 * packed at data + offsets[c] (mnw_encode_columns_dev's layout: offsets[c] = c * out_col_stride), with mins[c], bits[c];
 * its n values go to out_dev[c] (int64 or float32).  data (data_len readable bytes), offsets, mins, bits: DEVICE pointers; out_dev: a HOST array of
 * ncols DEVICE pointers.  jitter: CENTER or HASH (column c hashes with block id jitter->block_id0 + c).  Enqueued on the
 * context's stream without synchronising. */
MNW_API int mnw_decode_columns_dev(mnw_ctx *ctx, int64_t ncols, const mnw_column *cols, const uint8_t *data, int64_t data_len,
                                   const int64_t *offsets, const int64_t *mins, const int64_t *bits, int64_t n,
                                   const mnw_jitter *jitter, void *const *out_dev);

/* nfiles cubes back to back in aos (each nfile^3 particles); outputs are
 * [nfiles][3*subcells^3]; file f / axis k bytes at out + (3*f + k)*out_axis_stride,
 * lengths in out_len[3*f + k].  desc (HOST) holds 3 entries shared by all files
 * (desc_per_file = 0; periodic positions) or 3 per file (desc_per_file = 1;
 * non-periodic fields get their own limits per file, go/minp/minp.go:92-95). */
MNW_API int mnw_encode_vec3_subcells_dev(mnw_ctx *ctx, const mnw_float_desc *desc, int desc_per_file,
                                         const float *aos, int64_t nfile, int64_t subcells, int64_t nfiles,
                                         int64_t *mins, int64_t *bits, int64_t *offsets,
                                         uint8_t *out, int64_t out_axis_stride, int64_t *out_len);
MNW_API int mnw_decode_vec3_subcells_dev(mnw_ctx *ctx, const mnw_float_desc *desc, int desc_per_file,
                                         const uint8_t *data, int64_t data_axis_stride, const int64_t *offsets,
                                         const int64_t *mins, const int64_t *bits, int64_t nfile,
                                         int64_t subcells, int64_t nfiles, float wrap_L,
                                         const mnw_jitter *jitter, float *aos_out);

/* The whole of minp.Writer.Vectors / Reader.Vectors for nfiles device-resident cubes with NO host round trip inside:
 * for a non-periodic field bounds() (go/minp/minp.go:291-300), Nextafter32, the pixel counts (go/writer.go:73) and the
 * derived group constants are computed by kernels on the stream, so the call only enqueues.  desc_dev is a DEVICE array of
 * 3 * nfiles mnw_float_desc: written by the encode (what the files' group tails need; copy it back after mnw_sync),
 * read by the decode.  Other arguments as in mnw_encode_vec3_subcells_dev / mnw_decode_vec3_subcells_dev.  If a derived
 * group lies outside what the fused kernels cover (pixels < 1 or > 2^22), the generic kernels encode the batch. */
MNW_API int mnw_minp_encode_vectors_dev(mnw_ctx *ctx, const float *aos, int64_t nfile, int64_t subcells, int64_t nfiles,
                                        int periodic, float L, float dx, mnw_float_desc *desc_dev, int64_t *mins,
                                        int64_t *bits, int64_t *offsets, uint8_t *out, int64_t out_axis_stride,
                                        int64_t *out_len);
MNW_API int mnw_minp_decode_vectors_dev(mnw_ctx *ctx, const mnw_float_desc *desc_dev, const uint8_t *data,
                                        int64_t data_axis_stride, const int64_t *offsets, const int64_t *mins,
                                        const int64_t *bits, int64_t nfile, int64_t subcells, int64_t nfiles, int periodic,
                                        float L, const mnw_jitter *jitter, float *aos_out);

/* minp.Writer.Vectors limits for NON-periodic fields (go/minp/minp.go:92-95):
 * lo = min over particles, hi = Nextafter32(max, 2*max), per axis.  aos holds
 * nfiles cubes of np particles each; lo/hi are HOST arrays [nfiles][3].  The
 * _dev variant takes a DEVICE aos pointer (it synchronises to return lo/hi).
 * NaNs are ignored; if both -0 and +0 are extreme values the sign of the
 * returned zero may differ from Go's first-seen rule. */
MNW_API int mnw_vec3_limits(mnw_ctx *ctx, const float *aos, int64_t np, int64_t nfiles, float *lo, float *hi);
MNW_API int mnw_vec3_limits_dev(mnw_ctx *ctx, const float *aos, int64_t np, int64_t nfiles, float *lo, float *hi);

/* Device-pointer form of mnw_scan_offsets (sizes, offsets [nblocks], total [1]). */
MNW_API int mnw_scan_offsets_dev(mnw_ctx *ctx, const int64_t *nbytes, int64_t nblocks, int64_t base,
                                 int64_t *offsets, int64_t *total);

/* ---- pipelined host <-> device streaming (the staging either side of the kernels) ----------------------------------
 * Replaces the one-block-at-a-time file cursor of go/writer.go:90-141 / go/reader.go:114-127 / go/bit/bit.go:161-181 on
 * the host side of the boundary.  A pipe is a ring of `depth` slots (a private stream and device staging each).
 * Submitting a file ENQUEUES its upload, kernels and download and returns a ticket at once; with one host thread the
 * upload of file i+1, the kernels of file i and the download of file i-1 overlap.  The caller's buffers (pinned for
 * full speed, any host memory for correctness) must stay valid until mnw_pipe_wait(ticket) returns; results and errors
 * of a ticket are delivered by that wait.  Submitting blocks only when all `depth` slots are in flight.
 * Arguments as in mnw_minp_encode_vectors / mnw_decode_vec3_subcells.  Not thread-safe, like a context. */
typedef struct mnw_pipe mnw_pipe;
MNW_API int mnw_pipe_create(int device, int depth, mnw_pipe **out);
MNW_API void mnw_pipe_destroy(mnw_pipe *pipe);
MNW_API const char *mnw_pipe_last_error(const mnw_pipe *pipe);
MNW_API int mnw_pipe_minp_encode_vectors(mnw_pipe *pipe, const float *aos, int64_t nfile, int64_t subcells, int periodic,
                                         float L, float dx, mnw_float_desc desc_out[3], int64_t *mins, int64_t *bits,
                                         int64_t *offsets, uint8_t *out, int64_t out_axis_stride, int64_t out_len[3],
                                         int64_t *ticket);
MNW_API int mnw_pipe_minp_decode_vectors(mnw_pipe *pipe, const mnw_float_desc desc[3], const uint8_t *const data[3],
                                         const int64_t data_len[3], const int64_t *offsets, const int64_t *mins,
                                         const int64_t *bits, int64_t nfile, int64_t subcells, float wrap_L,
                                         const mnw_jitter *jitter, float *aos_out, int64_t *ticket);
MNW_API int mnw_pipe_poll(mnw_pipe *pipe);                  /* advance what can be advanced, never blocks */
MNW_API int mnw_pipe_wait(mnw_pipe *pipe, int64_t ticket);  /* that ticket is complete; returns its status */
MNW_API int mnw_pipe_drain(mnw_pipe *pipe);                 /* everything submitted is complete */

/* ---- multi-GPU: one process (or thread) and one context per GPU --------------------------------------------------------
 * Blocks are independent given the group parameters, so a file's blocks are split into contiguous ranges over the ranks;
 * every rank packs with rank-LOCAL offsets.  The only exchange is an NCCL all-gather of the per-block packed sizes
 * (NVLink / NVSwitch); the same scan on every rank then yields the file's global block offsets (blockIndex,
 * go/block_index.go:16-35) and base_r = offsets[first block of rank r], where the rank's bytes go (pwrite at
 * groupOffset + base_r).  Payload never crosses the interconnect.
 *   mnw_comm_unique_id   rank 0 creates the NCCL id (128 bytes) and hands it to the other ranks by its own means
 *   mnw_comm_init        every rank: joins the communicator on its context's device (collective)
 *   mnw_allgather_sizes  local_dev [count] int64 -> all_dev [count * nranks], DEVICE pointers, on the context's stream
 *   mnw_sharded_offsets_dev  the all-gather + the exclusive scan in one call: all_nbytes_dev / all_offsets_dev
 *                        [count * nranks], total_dev [1]
 * NCCL is loaded at run time (libnccl.so.2, or MNW_NCCL_LIB); a context without a communicator is a world of one. */
typedef struct { char bytes[128]; } mnw_comm_id;
MNW_API int mnw_comm_unique_id(mnw_comm_id *out);
MNW_API int mnw_comm_init(mnw_ctx *ctx, const mnw_comm_id *id, int nranks, int rank);
MNW_API int mnw_comm_destroy(mnw_ctx *ctx);
MNW_API int mnw_comm_size(const mnw_ctx *ctx);
MNW_API int mnw_comm_rank(const mnw_ctx *ctx);
MNW_API int mnw_allgather_sizes(mnw_ctx *ctx, const int64_t *local_dev, int64_t count, int64_t *all_dev);
MNW_API int mnw_sharded_offsets_dev(mnw_ctx *ctx, const int64_t *local_nbytes_dev, int64_t count, int64_t *all_nbytes_dev,
                                    int64_t *all_offsets_dev, int64_t *total_dev);

/* Per-kernel timing with CUDA events on the context's stream.  mnw_profile(ctx, 1)
 * starts recording the library's bandwidth-carrying kernels, mnw_profile(ctx, 0)
 * stops; mnw_profile_summary synchronises and writes a JSON array
 * [{"kernel": name, "launches": n, "ms": total}, ...] into buf and clears the log. */
MNW_API int mnw_profile(mnw_ctx *ctx, int on);
MNW_API int mnw_profile_summary(mnw_ctx *ctx, char *buf, int64_t cap);

/* Diagnostic: the fused kernels replace the IEEE float32 divide of go/group.go:319 by a
 * multiply with RN(1/dx) and two FMA corrections, and use the result only when the pixel
 * index lands in [1, pixels).  This runs both on every float32 bit pattern in
 * [first_bits, first_bits + count) and reports how many accepted results differ from the
 * divide (must be 0) and how many were accepted. */
MNW_API int mnw_selftest_fastdiv(mnw_ctx *ctx, const mnw_float_desc *desc, uint32_t first_bits, uint64_t count,
                                 uint64_t *mismatches, uint64_t *accepted);

/* Diagnostic: minh Log columns take float32(log10(float64 x)) (go/minh/minh.go:143) through a table + series form with
 * an exact fallback (device_math.cuh go_log10_f32).  This compares it with the restated Go algorithm on every float32
 * bit pattern in [first_bits, first_bits + count) and reports how many results differ (must be 0). */
MNW_API int mnw_selftest_log10(mnw_ctx *ctx, uint32_t first_bits, uint64_t count, uint64_t *mismatches);
/* The same for the read side's float32(pow(10, float64 x)): exp2 form with an exact fallback against the restated Go
 * math.Pow, on every float32 bit pattern in the range (must be 0). */
MNW_API int mnw_selftest_pow10(mnw_ctx *ctx, uint32_t first_bits, uint64_t count, uint64_t *mismatches);

/* out[i] = float32(math.Pow(10, float64(x[i]))): the read side of a minh Log column that is stored raw (Float32Group),
 * go/minh/minh.go:315-319.  FloatGroup Log columns get it inside mnw_decode_float_blocks (desc.log10).  HOST pointers.
 * Go's math.Pow / math.Exp are restated from the portable pure-Go sources (parity unpinned: the reference's own test of
 * this step is a tolerance test, go/minh/minh_test.go:110-113). */
MNW_API int mnw_pow10_f32(mnw_ctx *ctx, const float *x, int64_t n, float *out);

/* Which device path the last encode on ctx took: 0 = generic two-pass, 1 = fused single-read minp kernels
 * (k_pipe_vec3 / k_fused_vec3), 2 = fused group kernel (k_group_fused).  For tests and the benchmark. */
MNW_API int mnw_last_path(const mnw_ctx *ctx);
/* Force the generic path (testing both paths against the oracle). */
MNW_API void mnw_force_generic(mnw_ctx *ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* MINNOW_CUDA_H */
