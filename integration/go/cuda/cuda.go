// Package cuda binds libminnow_b200.so (include/minnow_cuda.h) for the reference's Go packages.
//
// This file is the reference-side glue a maintainer adds next to go/bit, go/group.go, go/minh and go/minp; the
// exported Go API of those packages and the bytes on disk do not change.  It is NOT built in this repository's
// image (no Go toolchain there: `go version` fails on both the build container and the GPU box, see DESIGN.md);
// the same calls are exercised through the ctypes binding minnow_b200/capi.py by tests/.
//
// Error convention: the reference panics (go/writer.go:34,92; go/bit/bit.go:31,86), so every wrapper re-panics with
// mnw_last_error.  Threading: a Ctx, like a minnow.Writer, is not safe for concurrent use -- one per goroutine
// (runtime.LockOSThread is not needed: every entry point sets its CUDA device itself).
package cuda

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../minnow_b200 -lminnow_b200 -Wl,-rpath,${SRCDIR}/../../../minnow_b200
#include <stdlib.h>
#include "minnow_cuda.h"
*/
import "C"

import (
	"runtime"
	"unsafe"
)

// Ctx is one mnw_ctx: a CUDA stream plus grow-only device scratch (the ArrayBuffer of go/bit/bit.go:146-206 on the GPU).
type Ctx struct{ h *C.mnw_ctx }

func New(device int) *Ctx {
	var h *C.mnw_ctx
	if rc := C.mnw_create(C.int(device), &h); rc != 0 {
		panic(C.GoString(C.mnw_last_error(nil)))
	}
	c := &Ctx{h}
	runtime.SetFinalizer(c, func(c *Ctx) { C.mnw_destroy(c.h) })
	return c
}

func (c *Ctx) check(rc C.int) {
	if rc != 0 {
		panic(C.GoString(C.mnw_last_error(c.h)))
	}
}

// FloatDesc mirrors floatGroup's parameters (go/group.go:265-276) + minh's column transform (go/minh/minh.go:141-149).
type FloatDesc struct {
	Low, High float32
	Pixels    int64
	Log, Clamp bool
}

func (d FloatDesc) c() C.mnw_float_desc {
	var o C.mnw_float_desc
	o.low, o.high, o.pixels, o.periodic = C.float(d.Low), C.float(d.High), C.int64_t(d.Pixels), 1 // go/writer.go:74
	if d.Log {
		o.log10 = 1
	}
	if d.Clamp {
		o.clamp = 1
	}
	return o
}

// Pixels = go/writer.go:73.
func Pixels(lo, hi, dx float32) int64 { return int64(C.mnw_float_group_pixels(C.float(lo), C.float(hi), C.float(dx))) }

// Block metadata of nblocks encoded blocks: what writeData appends to g.mins / g.bits and what blockIndex derives.
type Blocks struct {
	Mins, Bits, Offsets []int64
	Data                []byte
}

func newBlocks(nblocks, maxBytes int) Blocks {
	return Blocks{make([]int64, nblocks), make([]int64, nblocks), make([]int64, nblocks), make([]byte, maxBytes)}
}

func i64p(x []int64) *C.int64_t {
	if len(x) == 0 {
		return nil
	}
	return (*C.int64_t)(unsafe.Pointer(&x[0]))
}
func u8p(x []byte) *C.uint8_t {
	if len(x) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&x[0]))
}

// EncodeInts replaces nblocks calls of intGroup.writeData (go/group.go:242-255): x holds nblocks blocks of n values.
func (c *Ctx) EncodeInts(x []int64, n, nblocks int) Blocks {
	b := newBlocks(nblocks, 8*len(x)+8)
	var ln C.int64_t
	c.check(C.mnw_encode_int_group(c.h, i64p(x), C.int64_t(n), C.int64_t(nblocks), nil, i64p(b.Mins), i64p(b.Bits),
		i64p(b.Offsets), u8p(b.Data), C.int64_t(len(b.Data)), &ln))
	b.Data = b.Data[:ln]
	return b
}

// EncodeFloats replaces nblocks calls of floatGroup.writeData (go/group.go:312-327).
func (c *Ctx) EncodeFloats(d FloatDesc, x []float32, n, nblocks int) Blocks {
	b := newBlocks(nblocks, 8*len(x)+8)
	cd := d.c()
	var ln C.int64_t
	c.check(C.mnw_encode_float_group(c.h, &cd, (*C.float)(unsafe.Pointer(&x[0])), C.int64_t(n), C.int64_t(nblocks), nil,
		i64p(b.Mins), i64p(b.Bits), i64p(b.Offsets), u8p(b.Data), C.int64_t(len(b.Data)), &ln))
	b.Data = b.Data[:ln]
	return b
}

// DecodeInts replaces intGroup.readData (go/group.go:257-263) for the selected blocks of a group.
func (c *Ctx) DecodeInts(data []byte, b Blocks, n int, sel []int64, out []int64) {
	nsel := len(sel)
	if sel == nil {
		nsel = len(b.Mins)
	}
	c.check(C.mnw_decode_int_blocks(c.h, u8p(data), C.int64_t(len(data)), i64p(b.Offsets), i64p(b.Mins), i64p(b.Bits),
		C.int64_t(n), C.int64_t(nsel), i64p(sel), i64p(out)))
}

// Jitter: floatGroup.readData draws rand.Float64() per value (go/group.go:308); the stream is explicit here.
type Jitter struct {
	Mode int // 0 centre, 1 keyed hash, 2 caller's stream
	Seed uint64
	U    []float64
}

func (j Jitter) c() C.mnw_jitter {
	var o C.mnw_jitter
	o.mode, o.seed = C.int32_t(j.Mode), C.uint64_t(j.Seed)
	if len(j.U) > 0 {
		o.u_stream = (*C.double)(unsafe.Pointer(&j.U[0]))
	}
	return o
}

// DecodeFloats replaces floatGroup.readData (go/group.go:299-310); d.Log adds minh's float32(math.Pow(10, x)).
func (c *Ctx) DecodeFloats(d FloatDesc, data []byte, b Blocks, n int, sel []int64, j Jitter, out []float32) {
	nsel := len(sel)
	if sel == nil {
		nsel = len(b.Mins)
	}
	cd, cj := d.c(), j.c()
	c.check(C.mnw_decode_float_blocks(c.h, &cd, u8p(data), C.int64_t(len(data)), i64p(b.Offsets), i64p(b.Mins), i64p(b.Bits),
		C.int64_t(n), C.int64_t(nsel), i64p(sel), &cj, (*C.float)(unsafe.Pointer(&out[0]))))
}

// Vectors is the whole of minp.Writer.Vectors for one file (go/minp/minp.go:86-119): limits, pixels, sub-cell gather and
// the three FloatGroups; the caller appends desc / mins / bits to its group tails and writes data[k][:lens[k]].
func (c *Ctx) Vectors(vec [][3]float32, nFile, subCells int, periodic bool, L, dx float32) (desc [3]FloatDesc, b Blocks, lens [3]int64, stride int) {
	nb := 3 * subCells * subCells * subCells
	stride = 8*len(vec) + 8
	b = newBlocks(nb, 3*stride)
	var cd [3]C.mnw_float_desc
	per := C.int(0)
	if periodic {
		per = 1
	}
	c.check(C.mnw_minp_encode_vectors(c.h, (*C.float)(unsafe.Pointer(&vec[0][0])), C.int64_t(nFile), C.int64_t(subCells), per,
		C.float(L), C.float(dx), &cd[0], i64p(b.Mins), i64p(b.Bits), i64p(b.Offsets), u8p(b.Data), C.int64_t(stride),
		(*C.int64_t)(unsafe.Pointer(&lens[0]))))
	for k := 0; k < 3; k++ {
		desc[k] = FloatDesc{Low: float32(cd[k].low), High: float32(cd[k].high), Pixels: int64(cd[k].pixels)}
	}
	return
}

// ReadVectors is the loop body of minp.Reader.Vectors (go/minp/minp.go:191-206).
func (c *Ctx) ReadVectors(desc [3]FloatDesc, data [3][]byte, b Blocks, nFile, subCells int, wrapL float32, j Jitter, out [][3]float32) {
	var cd [3]C.mnw_float_desc
	var ptr [3]*C.uint8_t
	var ln [3]C.int64_t
	for k := 0; k < 3; k++ {
		cd[k], ptr[k], ln[k] = desc[k].c(), u8p(data[k]), C.int64_t(len(data[k]))
	}
	cj := j.c()
	c.check(C.mnw_decode_vec3_subcells(c.h, &cd[0], &ptr[0], &ln[0], i64p(b.Offsets), i64p(b.Mins), i64p(b.Bits),
		C.int64_t(nFile), C.int64_t(subCells), C.float(wrapL), &cj, (*C.float)(unsafe.Pointer(&out[0][0]))))
}

// Column is one quantised column of a minh block: Ints or Floats (with Desc) set.
type Column struct {
	Ints   []int64
	Floats []float32
	Desc   FloatDesc
}

// EncodeColumns is the per-column loop of minh.Writer.Block (go/minh/minh.go:99-139) in ONE call.
func (c *Ctx) EncodeColumns(cols []Column, n int) (mins, bits, nbytes []int64, out []byte, stride int) {
	nc := len(cols)
	cc := make([]C.mnw_column, nc)
	// the pointer table lives in C memory: cgo forbids Go pointers to Go pointers
	tab := (*[1 << 20]unsafe.Pointer)(C.malloc(C.size_t(nc) * C.size_t(unsafe.Sizeof(uintptr(0)))))[:nc:nc]
	defer C.free(unsafe.Pointer(&tab[0]))
	for i, col := range cols {
		if col.Floats != nil {
			cc[i].is_float, cc[i].desc = 1, col.Desc.c()
			tab[i] = unsafe.Pointer(&col.Floats[0])
		} else {
			tab[i] = unsafe.Pointer(&col.Ints[0])
		}
	}
	stride = 8*n + 16
	mins, bits, nbytes, out = make([]int64, nc), make([]int64, nc), make([]int64, nc), make([]byte, nc*stride)
	c.check(C.mnw_encode_columns(c.h, C.int64_t(nc), &cc[0], (*unsafe.Pointer)(unsafe.Pointer(&tab[0])), C.int64_t(n),
		i64p(mins), i64p(bits), i64p(nbytes), u8p(out), C.int64_t(stride)))
	return
}

// Boundary is BoundaryWriter.Coordinates (go/minh/boundary.go:39-51): the per-cell index lists stay on the GPU.
func (c *Ctx) Boundary(x, y, z []float32, L, boundary float32, cells int) (sizes []int64, total int64) {
	sizes = make([]int64, cells*cells*cells)
	var t C.int64_t
	c.check(C.mnw_boundary_coordinates(c.h, (*C.float)(unsafe.Pointer(&x[0])), (*C.float)(unsafe.Pointer(&y[0])),
		(*C.float)(unsafe.Pointer(&z[0])), C.int64_t(len(x)), C.float(L), C.float(boundary), C.int64_t(cells), i64p(sizes), &t))
	return sizes, int64(t)
}

// BoundaryInts / BoundaryFloats are the IntGroup / FloatGroup branches of BoundaryWriter.Column (go/minh/boundary.go:184-225).
func (c *Ctx) BoundaryInts(col []int64, cells int, total int64) Blocks {
	b := newBlocks(cells*cells*cells, int(8*total)+64)
	var ln C.int64_t
	c.check(C.mnw_boundary_encode_int_column(c.h, i64p(col), C.int64_t(len(col)), i64p(b.Mins), i64p(b.Bits), i64p(b.Offsets),
		u8p(b.Data), C.int64_t(len(b.Data)), &ln))
	b.Data = b.Data[:ln]
	return b
}

// ParseText is the body of text.Reader.Block (go/text/text.go:181-200) for one block of bytes.
func (c *Ctx) ParseText(buf []byte, sep, comment byte, icols, fcols []int32) (ints [][]int64, floats [][]float32) {
	var rows, nfb C.int64_t
	c.check(C.mnw_text_parse_block(c.h, (*C.char)(unsafe.Pointer(&buf[0])), C.int64_t(len(buf)), C.char(sep), C.char(comment),
		C.int(len(icols)), (*C.int)(unsafe.Pointer(&icols[0])), C.int(len(fcols)), (*C.int)(unsafe.Pointer(&fcols[0])), &rows, &nfb))
	iflat, fflat, fb := make([]int64, len(icols)*int(rows)), make([]float32, len(fcols)*int(rows)), make([]int64, 3*int(nfb)+1)
	c.check(C.mnw_text_columns(c.h, i64p(iflat), (*C.float)(unsafe.Pointer(&fflat[0])), i64p(fb)))
	for k := 0; k < int(nfb); k++ { // the fields Eisel-Lemire could not decide: strconv itself
		off, ln := fb[3*k+2]&(1<<40-1), fb[3*k+2]>>40
		fflat[int(fb[3*k+1])*int(rows)+int(fb[3*k])] = float32(mustParseFloat(string(buf[off : off+ln])))
	}
	for j := range icols {
		ints = append(ints, iflat[j*int(rows):(j+1)*int(rows)])
	}
	for j := range fcols {
		floats = append(floats, fflat[j*int(rows):(j+1)*int(rows)])
	}
	return
}

// Pipe streams minp files through a ring of slots with ONE goroutine (upload, kernels and download of different files
// overlap): the staging of go/writer.go:107-141 / go/reader.go:114-127 in pipelined form.
type Pipe struct{ h *C.mnw_pipe }

func NewPipe(device, depth int) *Pipe {
	var h *C.mnw_pipe
	if rc := C.mnw_pipe_create(C.int(device), C.int(depth), &h); rc != 0 {
		panic(C.GoString(C.mnw_last_error(nil)))
	}
	p := &Pipe{h}
	runtime.SetFinalizer(p, func(p *Pipe) { C.mnw_pipe_destroy(p.h) })
	return p
}

// Wait blocks until the ticket's results are in the buffers given at submission.
func (p *Pipe) Wait(ticket int64) {
	if rc := C.mnw_pipe_wait(p.h, C.int64_t(ticket)); rc != 0 {
		panic(C.GoString(C.mnw_pipe_last_error(p.h)))
	}
}

// Comm: the one exchange of a snapshot sharded by block range over GPUs (go/block_index.go applied to the whole file).
func UniqueID() (id [128]byte) {
	if rc := C.mnw_comm_unique_id((*C.mnw_comm_id)(unsafe.Pointer(&id[0]))); rc != 0 {
		panic(C.GoString(C.mnw_last_error(nil)))
	}
	return
}
func (c *Ctx) CommInit(id [128]byte, nranks, rank int) {
	c.check(C.mnw_comm_init(c.h, (*C.mnw_comm_id)(unsafe.Pointer(&id[0])), C.int(nranks), C.int(rank)))
}
