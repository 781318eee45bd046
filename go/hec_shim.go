// hec_shim.go -- cgo binding of libhec (include/hec.h) for the reference's Go host
// (dwkim606/optimal_conv: main.go, conv.go, eval.go).  Drop this file next to conv.go (package main).
//
// WRITE-ONLY in this repository: the build image has no Go toolchain, so the file is never compiled
// here.  It is kept mechanical -- every C.hec_* call below is checked by name and argument count
// against include/hec.h in tests/test_go_shim.py, and the same calls are exercised through the ctypes
// binding (optimal_conv_b200/hec.py) by the GPU tests.
//
// Two granularities (INTEGRATION.md):
//   1. coarse: conv_then_pack_gpu replaces the body of conv_then_pack (conv.go:522-546) + the bias Add
//      (eval.go:258); prepKerGPU replaces the plaintext loop of prep_Ker (conv.go:510-515);
//   2. fine:   gpuEvaluator embeds the Lattigo evaluator and overrides the methods the conv path calls
//      on context.pack_evaluator / context.evaluator (main.go:39-40).
package main

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -L${SRCDIR}/../optimal_conv_b200 -lhec
#include "hec.h"
#include <stdlib.h>
*/
import "C"

import (
	"runtime"
	"unsafe"

	"github.com/dwkim606/test_lattigo/ckks"
	"github.com/dwkim606/test_lattigo/ring"
	"github.com/dwkim606/test_lattigo/rlwe"
)

// gpuEval owns one libhec context = one evaluator bound to one GPU (ckks.NewEvaluator, conv.go:258).
type gpuEval struct {
	ctx    *C.hec_ctx
	params ckks.Parameters
	pts    map[*ckks.Plaintext]*C.hec_pt // plaintexts uploaded once per layer and reused across images
}

// chk keeps the reference's convention: panic on any inconsistency (conv.go:541-543, eval.go:252-257).
func (g *gpuEval) chk(rc C.int) {
	if rc != 0 {
		panic(C.GoString(C.hec_last_error(g.ctx)))
	}
}

// limbPtrs returns the C array {&Coeffs[0][0], ..., &Coeffs[n-1][0]}.  The array itself is C memory, so no
// Go pointer to a Go pointer crosses the boundary (cgo rule); libhec copies before returning.
func limbPtrs(p *ring.Poly, n int) (**C.uint64_t, func()) {
	arr := (*[1 << 16]*C.uint64_t)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	for i := 0; i < n; i++ {
		arr[i] = (*C.uint64_t)(unsafe.Pointer(&p.Coeffs[i][0]))
	}
	return (**C.uint64_t)(unsafe.Pointer(arr)), func() { C.free(unsafe.Pointer(arr)) }
}

// newGpuEval replaces ckks.NewEvaluator(new_params, rlwe.EvaluationKey{Rlk: rlk, Rtks: rtks}) (conv.go:258, main.go:430-441).
func newGpuEval(params ckks.Parameters, rlk *rlwe.RelinearizationKey, rtks *rlwe.RotationKeySet, maxLevel, device int) *gpuEval {
	runtime.LockOSThread() // a context is not re-entrant, like a Lattigo evaluator
	g := &gpuEval{params: params, pts: map[*ckks.Plaintext]*C.hec_pt{}}
	Q, P := params.Q(), params.P()
	if rc := C.hec_ctx_create(&g.ctx, C.int(params.LogN()), (*C.uint64_t)(&Q[0]), C.int(len(Q)),
		(*C.uint64_t)(&P[0]), C.int(len(P)), C.int(device)); rc != 0 {
		panic("hec_ctx_create failed: no CUDA device or unsupported parameters")
	}
	full := len(Q) + len(P)
	keyPtrs := func(swk *rlwe.SwitchingKey) (**C.uint64_t, func()) { // rlwe.SwitchingKey{Value [][2]*ring.Poly}
		arr := (*[1 << 20]*C.uint64_t)(C.malloc(C.size_t(len(swk.Value)*2*full) * 8))
		for d := range swk.Value {
			for k := 0; k < 2; k++ {
				for t := 0; t < full; t++ {
					arr[(d*2+k)*full+t] = (*C.uint64_t)(unsafe.Pointer(&swk.Value[d][k].Coeffs[t][0]))
				}
			}
		}
		return (**C.uint64_t)(unsafe.Pointer(arr)), func() { C.free(unsafe.Pointer(arr)) }
	}
	if rtks != nil {
		for galEl, swk := range rtks.Keys {
			p, free := keyPtrs(swk)
			g.chk(C.hec_swk_upload(g.ctx, C.uint64_t(galEl), C.int(maxLevel), p))
			free()
		}
	}
	if rlk != nil {
		p, free := keyPtrs(rlk.Keys[0])
		g.chk(C.hec_rlk_upload(g.ctx, C.int(maxLevel), p))
		free()
	}
	return g
}

func (g *gpuEval) close() { C.hec_ctx_destroy(g.ctx) }

func (g *gpuEval) uploadPt(pt *ckks.Plaintext) *C.hec_pt {
	if h, ok := g.pts[pt]; ok {
		return h
	}
	var h *C.hec_pt
	p, free := limbPtrs(pt.Value, pt.Level()+1)
	defer free()
	g.chk(C.hec_pt_upload(g.ctx, C.int(pt.Level()), p, C.double(pt.Scale), &h))
	g.pts[pt] = h
	return h
}

// dropPt forgets a plaintext the host is done with (end of a layer); cached conv plans that read it go with it.
func (g *gpuEval) dropPt(pt *ckks.Plaintext) {
	if h, ok := g.pts[pt]; ok {
		C.hec_pt_free(g.ctx, h)
		delete(g.pts, pt)
	}
}

func (g *gpuEval) uploadCt(ct *ckks.Ciphertext) *C.hec_ct {
	var h *C.hec_ct
	c0, f0 := limbPtrs(ct.Value[0], ct.Level()+1)
	c1, f1 := limbPtrs(ct.Value[1], ct.Level()+1)
	defer f0()
	defer f1()
	g.chk(C.hec_ct_upload(g.ctx, C.int(ct.Level()), c0, c1, C.double(ct.Scale), &h))
	return h
}

// downloadCt copies a device ciphertext into a fresh ckks.Ciphertext and releases the device handle.
func (g *gpuEval) downloadCt(h *C.hec_ct) *ckks.Ciphertext {
	level := int(C.hec_ct_level(h))
	res := ckks.NewCiphertext(g.params, 1, level, float64(C.hec_ct_scale(h)))
	o0, f0 := limbPtrs(res.Value[0], level+1)
	o1, f1 := limbPtrs(res.Value[1], level+1)
	defer f0()
	defer f1()
	g.chk(C.hec_ct_download(g.ctx, h, o0, o1))
	C.hec_ct_free(g.ctx, h)
	return res
}

// pinPoly page-locks the backing arrays of a polynomial (Go's heap does not move objects) so that uploads,
// downloads and hec_plan_submit_host copy asynchronously at full PCIe rate.
func (g *gpuEval) pinPoly(p *ring.Poly) {
	for i := range p.Coeffs {
		g.chk(C.hec_host_register(g.ctx, unsafe.Pointer(&p.Coeffs[i][0]), C.size_t(8*len(p.Coeffs[i]))))
	}
}

// ---------------------------------------------------------------------------------------------
// 1. coarse entry points
// ---------------------------------------------------------------------------------------------

// prepKerGPU replaces the plaintext loop at the end of prep_Ker (conv.go:510-515): the float reshaping above it
// stays as it is, EncodeCoeffs + ToNTT of all max_bat vectors run on the device in one call.
func prepKerGPU(g *gpuEval, max_ker_rs [][]float64, pos, in_wid, max_bat, ker_wid, ECD_LV int) []*C.hec_pt {
	n := g.params.N()
	flat := make([]float64, max_bat*n) // [max_bat][N], back to back
	for i := 0; i < max_bat; i++ {
		copy(flat[i*n:], encode_ker_final(max_ker_rs, pos, i, in_wid, max_bat, ker_wid))
	}
	out := make([]*C.hec_pt, max_bat)
	g.chk(C.hec_encode_coeffs_many(g.ctx, (*C.double)(&flat[0]), C.int(max_bat), C.int(n), C.int(ECD_LV),
		C.double(g.params.Scale()), (**C.hec_pt)(unsafe.Pointer(&out[0]))))
	return out
}

// conv_then_pack_gpu is a drop-in for conv_then_pack (conv.go:522-546) [+ eval.go:252-258 when bias != nil].
// Plaintexts are uploaded on first use and stay resident (uploadPt), so the second image of a layer reuses them and
// the plan libhec cached for them (hec_plan_cache_size).
func conv_then_pack_gpu(g *gpuEval, ctxt_in *ckks.Ciphertext, pl_ker, plain_idx []*ckks.Plaintext,
	bias *ckks.Plaintext, max_ob, norm int, out_scale float64) *ckks.Ciphertext {
	in := g.uploadCt(ctxt_in)
	ker := (*[256]*C.hec_pt)(C.malloc(256 * 8))
	idx := (*[16]*C.hec_pt)(C.malloc(16 * 8))
	defer C.free(unsafe.Pointer(ker))
	defer C.free(unsafe.Pointer(idx))
	for i := 0; i < max_ob; i++ {
		ker[i] = nil
		if i%norm == 0 {
			ker[i] = g.uploadPt(pl_ker[i])
		}
	}
	for i := range plain_idx {
		idx[i] = g.uploadPt(plain_idx[i])
	}
	var b *C.hec_pt
	if bias != nil {
		b = g.uploadPt(bias)
	}
	var out *C.hec_ct
	g.chk(C.hec_conv_then_pack(g.ctx, in, (**C.hec_pt)(unsafe.Pointer(ker)), C.int(max_ob), C.int(norm),
		C.double(out_scale), (**C.hec_pt)(unsafe.Pointer(idx)), b, C.HEC_CONV_FUSED, &out))
	C.hec_ct_free(g.ctx, in)
	return g.downloadCt(out)
}

// conv_bl_gpu covers the timed interval of evalConv_BN_BL_test (eval.go:108-131): preConv_BL + the postConv_BL loop +
// RotateNew/Add + bias.  pl_taps[i*ker_wid*ker_wid+tap] are the plaintexts postConv_BL encodes (conv.go:165-166).
func conv_bl_gpu(g *gpuEval, ct_input *ckks.Ciphertext, in_wid, ker_wid, rot_iters, rot_step int,
	pl_taps []*ckks.Plaintext, pl_bn_b *ckks.Plaintext) *ckks.Ciphertext {
	in := g.uploadCt(ct_input)
	taps := (*[1 << 16]*C.hec_pt)(C.malloc(C.size_t(len(pl_taps)) * 8))
	defer C.free(unsafe.Pointer(taps))
	for i, pt := range pl_taps {
		taps[i] = g.uploadPt(pt)
	}
	var out *C.hec_ct
	g.chk(C.hec_conv_bl(g.ctx, in, C.int(in_wid), C.int(ker_wid), C.int(rot_iters), C.int(rot_step),
		(**C.hec_pt)(unsafe.Pointer(taps)), g.uploadPt(pl_bn_b), &out))
	C.hec_ct_free(g.ctx, in)
	return g.downloadCt(out)
}

// evalReLU_gpu replaces evalReLU (conv.go:435-480) followed by MulByPow2 (eval.go:476).
func evalReLU_gpu(g *gpuEval, ct *ckks.Ciphertext, alpha float64, pow int) *ckks.Ciphertext {
	in := g.uploadCt(ct)
	var out *C.hec_ct
	g.chk(C.hec_eval_relu(g.ctx, in, C.double(alpha), C.double(g.params.Scale()), &out))
	g.chk(C.hec_mul_by_pow2(g.ctx, out, C.int(pow)))
	C.hec_ct_free(g.ctx, in)
	return g.downloadCt(out)
}

// ---------------------------------------------------------------------------------------------
// 2. fine: a ckks.Evaluator whose hot methods run on the device
// ---------------------------------------------------------------------------------------------

// gpuEvaluator embeds the Lattigo evaluator -- the methods that are not on the conv path keep working on the
// host -- and overrides the ones conv.go / eval.go call.  Device handles live in a side table; a ciphertext is
// downloaded when the host asks for it (sync).
type gpuEvaluator struct {
	ckks.Evaluator
	g   *gpuEval
	dev map[*ckks.Ciphertext]*C.hec_ct
}

func (e *gpuEvaluator) h(ct *ckks.Ciphertext) *C.hec_ct {
	if d, ok := e.dev[ct]; ok {
		return d
	}
	d := e.g.uploadCt(ct)
	e.dev[ct] = d
	return d
}

// wrap registers a device result under a fresh host ciphertext whose limbs are filled by sync().
func (e *gpuEvaluator) wrap(d *C.hec_ct) *ckks.Ciphertext {
	ct := ckks.NewCiphertext(e.g.params, 1, int(C.hec_ct_level(d)), float64(C.hec_ct_scale(d)))
	e.dev[ct] = d
	return ct
}

// sync brings the host copy of ct up to date (call before Decrypt or any host-side evaluator method).
func (e *gpuEvaluator) sync(ct *ckks.Ciphertext) {
	d, ok := e.dev[ct]
	if !ok {
		return
	}
	level := int(C.hec_ct_level(d))
	ct.Value[0].Coeffs = ct.Value[0].Coeffs[:level+1]
	ct.Value[1].Coeffs = ct.Value[1].Coeffs[:level+1]
	ct.Scale = float64(C.hec_ct_scale(d))
	o0, f0 := limbPtrs(ct.Value[0], level+1)
	o1, f1 := limbPtrs(ct.Value[1], level+1)
	defer f0()
	defer f1()
	e.g.chk(C.hec_ct_download(e.g.ctx, d, o0, o1))
}

func (e *gpuEvaluator) MulNew(ct *ckks.Ciphertext, pt *ckks.Plaintext) *ckks.Ciphertext { // conv.go:168,288,527
	var out *C.hec_ct
	e.g.chk(C.hec_mul_pt_new(e.g.ctx, e.h(ct), e.g.uploadPt(pt), &out))
	return e.wrap(out)
}

func (e *gpuEvaluator) SetScale(ct *ckks.Ciphertext, scale float64) { // conv.go:528
	e.g.chk(C.hec_set_scale(e.g.ctx, e.h(ct), C.double(scale)))
}

func (e *gpuEvaluator) RescaleInPlace(ct *ckks.Ciphertext, minScale float64) {
	e.g.chk(C.hec_rescale(e.g.ctx, e.h(ct), C.double(minScale)))
}

func (e *gpuEvaluator) MultByConstInPlace(ct *ckks.Ciphertext, constant float64) {
	e.g.chk(C.hec_mult_by_const(e.g.ctx, e.h(ct), C.double(constant)))
}

func (e *gpuEvaluator) SubNew(a, b *ckks.Ciphertext) *ckks.Ciphertext { // conv.go:289
	var out *C.hec_ct
	e.g.chk(C.hec_sub_new(e.g.ctx, e.h(a), e.h(b), &out))
	return e.wrap(out)
}

func (e *gpuEvaluator) AddNew(a, b *ckks.Ciphertext) *ckks.Ciphertext {
	var out *C.hec_ct
	e.g.chk(C.hec_add_new(e.g.ctx, e.h(a), e.h(b), &out))
	return e.wrap(out)
}

func (e *gpuEvaluator) AddCt(a, b, out *ckks.Ciphertext) { // Add(ct, ct, ct): conv.go:171,290,292
	e.g.chk(C.hec_add(e.g.ctx, e.h(a), e.h(b), e.h(out)))
}

func (e *gpuEvaluator) SubCt(a, b, out *ckks.Ciphertext) {
	e.g.chk(C.hec_sub(e.g.ctx, e.h(a), e.h(b), e.h(out)))
}

func (e *gpuEvaluator) AddPt(ct *ckks.Ciphertext, pt *ckks.Plaintext) { // Add(ct, pt, ct): eval.go:130,258
	e.g.chk(C.hec_add_pt(e.g.ctx, e.h(ct), e.g.uploadPt(pt)))
}

func (e *gpuEvaluator) RotateGal(ct *ckks.Ciphertext, galEl uint64, out *ckks.Ciphertext) { // conv.go:291
	e.g.chk(C.hec_rotate_gal(e.g.ctx, e.h(ct), C.uint64_t(galEl), e.h(out)))
}

func (e *gpuEvaluator) RotateNew(ct *ckks.Ciphertext, k int) *ckks.Ciphertext { // eval.go:123
	var out *C.hec_ct
	e.g.chk(C.hec_rotate_new(e.g.ctx, e.h(ct), C.int(k), &out))
	return e.wrap(out)
}

func (e *gpuEvaluator) RotateHoisted(ct *ckks.Ciphertext, rotations []int) map[int]*ckks.Ciphertext { // conv.go:133
	n := len(rotations)
	rots := (*[1 << 16]C.int)(C.malloc(C.size_t(n) * 4))
	outs := (*[1 << 16]*C.hec_ct)(C.malloc(C.size_t(n) * 8))
	defer C.free(unsafe.Pointer(rots))
	defer C.free(unsafe.Pointer(outs))
	for i, r := range rotations {
		rots[i] = C.int(r)
	}
	e.g.chk(C.hec_rotate_hoisted(e.g.ctx, e.h(ct), (*C.int)(unsafe.Pointer(rots)), C.int(n), (**C.hec_ct)(unsafe.Pointer(outs))))
	res := make(map[int]*ckks.Ciphertext, n)
	for i, r := range rotations {
		res[r] = e.wrap(outs[i])
	}
	return res
}

func (e *gpuEvaluator) MulRelinNew(a, b *ckks.Ciphertext) *ckks.Ciphertext { // Mul + Relinearize, conv.go:473-474
	var out *C.hec_ct
	e.g.chk(C.hec_mul_relin_new(e.g.ctx, e.h(a), e.h(b), &out))
	return e.wrap(out)
}

func (e *gpuEvaluator) EvaluatePolyCoeffs(ct *ckks.Ciphertext, coeffs []float64, targetScale float64) *ckks.Ciphertext { // conv.go:460-471
	var out *C.hec_ct
	e.g.chk(C.hec_evaluate_poly(e.g.ctx, e.h(ct), (*C.double)(&coeffs[0]), C.int(len(coeffs)-1), C.double(targetScale),
		C.double(e.g.params.Scale()), &out))
	return e.wrap(out)
}

func (e *gpuEvaluator) AddConst(ct *ckks.Ciphertext, c float64) {
	e.g.chk(C.hec_add_const(e.g.ctx, e.h(ct), C.double(c)))
}

func (e *gpuEvaluator) DropLevel(ct *ckks.Ciphertext, levels int) {
	e.g.chk(C.hec_drop_level(e.g.ctx, e.h(ct), C.int(levels)))
}

func (e *gpuEvaluator) MulByPow2(ct *ckks.Ciphertext, pow2 int) { // eval.go:476
	e.g.chk(C.hec_mul_by_pow2(e.g.ctx, e.h(ct), C.int(pow2)))
}

func (e *gpuEvaluator) CopyNew(ct *ckks.Ciphertext) *ckks.Ciphertext {
	var out *C.hec_ct
	e.g.chk(C.hec_ct_copy_new(e.g.ctx, e.h(ct), &out))
	return e.wrap(out)
}

// release drops the device copy of a ciphertext the host is done with.
func (e *gpuEvaluator) release(ct *ckks.Ciphertext) {
	if d, ok := e.dev[ct]; ok {
		C.hec_ct_free(e.g.ctx, d)
		delete(e.dev, ct)
	}
}
