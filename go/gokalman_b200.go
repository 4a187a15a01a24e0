// Package gokalmanb200 is the cgo shim that puts the B200 engine (libgokalman_b200.so, C-ABI in
// include/gokalman_b200.h) behind gokalman's own interfaces.  SOURCE ONLY: the build image has no Go
// toolchain, so this file is never compiled or tested here; the Python mirror
// (gokalman_b200/api.py) is the tested caller of the same C-ABI.
//
// Build (on a machine with Go, CUDA 12.9 and the built library):
//   CGO_CFLAGS="-I${REPO}/include" CGO_LDFLAGS="-L${REPO}/gokalman_b200 -lgokalman_b200" go build
package gokalmanb200

/*
#include <stdlib.h>
#include "gokalman_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"unsafe"

	"github.com/ChristopherRabotin/gokalman"
	"github.com/gonum/matrix/mat64"
)

func raw(m mat64.Matrix) []float64 { // row-major copy, like mat64.DenseCopyOf(m).RawMatrix().Data
	r, c := m.Dims()
	out := make([]float64, r*c)
	for i := 0; i < r; i++ {
		for j := 0; j < c; j++ {
			out[i*c+j] = m.At(i, j)
		}
	}
	return out
}

func ptr(v []float64) *C.double {
	if len(v) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&v[0]))
}

func lastErr(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return errors.New(C.GoString(C.gkb_last_error()))
}

// Vanilla is a batch of N vanilla Kalman filters on one GPU; N = 1 implements gokalman.LDKF.
type Vanilla struct {
	h          *C.gkb_filter
	n, m, c, N int
	noise      gokalman.Noise
	F, G, H    mat64.Matrix
}

// NewVanilla mirrors gokalman.NewVanilla (vanilla.go:21-40) plus the batch size and device.
func NewVanilla(x0 *mat64.Vector, Covar0 mat64.Symmetric, F, G, H mat64.Matrix, noise gokalman.Noise, nFilters, device int) (*Vanilla, error) {
	n, _ := F.Dims()
	m, _ := H.Dims()
	_, c := G.Dims()
	kf := &Vanilla{n: n, m: m, c: c, N: nFilters, noise: noise, F: F, G: G, H: H}
	rc := C.gkb_create_lti(C.GKB_VANILLA, C.int(n), C.int(m), C.int(c), C.int64_t(nFilters), C.int(device),
		ptr(raw(x0)), 0, ptr(raw(Covar0)), ptr(raw(F)), ptr(raw(G)), ptr(raw(H)),
		ptr(raw(noise.ProcessMatrix())), ptr(raw(noise.MeasurementMatrix())), &kf.h)
	return kf, lastErr(rc)
}

// Update implements LDKF.Update(measurement, control) (kalman.go:36) for N = 1.
func (kf *Vanilla) Update(measurement, control *mat64.Vector) (gokalman.Estimate, error) {
	y, u := raw(measurement), raw(control)
	est := &Estimate{n: kf.n, m: kf.m,
		state: make([]float64, kf.n), meas: make([]float64, kf.m), innov: make([]float64, kf.m),
		covar: make([]float64, kf.n*kf.n), pred: make([]float64, kf.n*kf.n), gain: make([]float64, kf.n*kf.m)}
	var status C.int32_t
	out := C.gkb_outputs{mem: C.GKB_HOST, every_step: 0, state: ptr(est.state), meas: ptr(est.meas),
		innov: ptr(est.innov), covar: ptr(est.covar), pred_covar: ptr(est.pred), gain: ptr(est.gain), status: &status}
	if rc := C.gkb_update(kf.h, 1, ptr(y), 1, ptr(u), C.GKB_HOST, &out); rc != 0 {
		return nil, lastErr(rc)
	}
	if status != 0 {
		return nil, errors.New("could not invert `H*P_kp1_minus*H' + R`") // vanilla.go:164-167
	}
	return est, nil
}

// UpdateBatch is the new batched entry point: `steps` updates of all N filters in one launch.
// y is [steps][m][N] (filter index fastest), u is [steps][c]; outputs as requested in `out`.
func (kf *Vanilla) UpdateBatch(steps int, y, u []float64, out *C.gkb_outputs) error {
	return lastErr(C.gkb_update(kf.h, C.int(steps), ptr(y), 0, ptr(u), C.GKB_HOST, out))
}

func (kf *Vanilla) SetMeasurementMatrix(H mat64.Matrix) {
	m, _ := H.Dims()
	kf.H, kf.m = H, m
	C.gkb_set_measurement_matrix(kf.h, C.int(m), ptr(raw(H)))
}
func (kf *Vanilla) SetNoise(n gokalman.Noise) {
	r, _ := n.MeasurementMatrix().Dims()
	kf.noise = n
	C.gkb_set_noise(kf.h, ptr(raw(n.ProcessMatrix())), C.int(r), ptr(raw(n.MeasurementMatrix())))
}
func (kf *Vanilla) SetStateTransition(F mat64.Matrix) { kf.F = F; C.gkb_set_state_transition(kf.h, ptr(raw(F))) }
func (kf *Vanilla) SetInputControl(G mat64.Matrix) {
	_, c := G.Dims()
	kf.G, kf.c = G, c
	C.gkb_set_input_control(kf.h, C.int(c), ptr(raw(G)))
}
func (kf *Vanilla) GetNoise() gokalman.Noise            { return kf.noise }
func (kf *Vanilla) GetStateTransition() mat64.Matrix    { return kf.F }
func (kf *Vanilla) GetInputControl() mat64.Matrix       { return kf.G }
func (kf *Vanilla) GetMeasurementMatrix() mat64.Matrix  { return kf.H }
func (kf *Vanilla) Reset()                              { C.gkb_reset(kf.h); kf.noise.Reset() }
func (kf *Vanilla) String() string                      { return "gokalman_b200.Vanilla" }
func (kf *Vanilla) Close()                              { C.gkb_destroy(kf.h) }

// Estimate implements gokalman.Estimate (kalman.go:64-72) over the arrays gkb_update filled.
type Estimate struct {
	n, m                                   int
	state, meas, innov, covar, pred, gain []float64
}

func (e *Estimate) State() *mat64.Vector            { return mat64.NewVector(e.n, e.state) }
func (e *Estimate) Measurement() *mat64.Vector      { return mat64.NewVector(e.m, e.meas) }
func (e *Estimate) Innovation() *mat64.Vector       { return mat64.NewVector(e.m, e.innov) }
func (e *Estimate) Covariance() mat64.Symmetric     { return mat64.NewSymDense(e.n, e.covar) }
func (e *Estimate) PredCovariance() mat64.Symmetric { return mat64.NewSymDense(e.n, e.pred) }
func (e *Estimate) Gain() mat64.Matrix              { return mat64.NewDense(e.n, e.m, e.gain) }
func (e *Estimate) String() string                  { return "gokalman_b200.Estimate" }
func (e *Estimate) IsWithinNσ(N float64) bool { // vanilla.go:231-239
	for i := 0; i < e.n; i++ {
		b := N * sqrt(e.covar[i*e.n+i])
		if e.state[i] > b || e.state[i] < -b {
			return false
		}
	}
	return true
}

// ChiSquare runs NewMonteCarloRuns + NewChiSquare (montecarlo.go:92-119, chisquare.go:16-95) fused on
// the GPU and returns (NISmeans, NEESmeans) in the reference's order.
func ChiSquare(kind int, n, m, c int, F, G, H, Q, R, x0Truth, x0Filter, P0 []float64, trials int64, steps int,
	controls []float64, seed uint64, device int) (nis, nees []float64, err error) {
	nis, nees = make([]float64, steps), make([]float64, steps)
	cfg := C.gkb_mc_config{kind: C.int(kind), n: C.int(n), m: C.int(m), c: C.int(c), F: ptr(F), G: ptr(G), H: ptr(H),
		Q: ptr(Q), R: ptr(R), x0_truth: ptr(x0Truth), x0_filter: ptr(x0Filter), P0: ptr(P0), trials: C.int64_t(trials),
		steps: C.int(steps), controls: ptr(controls), noise_mode: C.GKB_NOISE_PHILOX, seed: C.uint64_t(seed),
		with_nees: 1, with_nis: 1, device: C.int(device)}
	var first C.int32_t
	out := C.gkb_mc_outputs{mem: C.GKB_HOST, nis: ptr(nis), nees: ptr(nees), first_error: &first}
	if rc := C.gkb_mc_chisquare(&cfg, &out); rc != 0 {
		return nil, nil, lastErr(rc)
	}
	if first != 0 {
		panic("a trial's Update failed") // chisquare.go:40-42 panics
	}
	return nis, nees, nil
}

func sqrt(x float64) float64 { // avoid importing math for one call in this sketch
	z := x
	for i := 0; i < 60 && z > 0; i++ {
		z = 0.5 * (z + x/z)
	}
	return z
}

// ---- entry points added after the first sketch (same caveat: source only) -------------------------------

// SmoothAll mirrors HybridKF.SmoothAll / SRIF.SmoothAll (hybrid.go:209-238, srif.go:165-192) on the stored
// histories of a batch: Phi [steps][n*n][N] (or [steps][n*n] when phiShared), state [steps][n][N],
// covar [steps][n*n][N], all overwritten in place for k < steps-1.
func SmoothAll(n, steps int, nFilters int64, device int, Phi []float64, phiShared bool, state, covar []float64) error {
	shared := C.int(0)
	if phiShared {
		shared = 1
	}
	status := make([]C.int32_t, nFilters)
	rc := C.gkb_smooth_all(C.int(n), C.int(steps), C.int64_t(nFilters), C.int(device), ptr(Phi), shared,
		ptr(state), ptr(covar), C.GKB_HOST, &status[0])
	if err := lastErr(rc); err != nil {
		return err
	}
	for _, s := range status {
		if s != 0 {
			return errors.New("provided STM Φ is not invertible") // hybrid.go:222
		}
	}
	return nil
}

// BatchSolve mirrors NewBatchKF + SetNextMeasurement x steps + Solve (batch.go:34-79) for N batch filters.
func BatchSolve(n, m, steps int, nFilters int64, device int, R, H []float64, hShared bool, realObs, computedObs []float64) (xHat0, P0 []float64, err error) {
	shared := C.int(0)
	if hShared {
		shared = 1
	}
	xHat0 = make([]float64, int64(n)*nFilters)
	P0 = make([]float64, int64(n*n)*nFilters)
	rc := C.gkb_batch_solve(C.int(n), C.int(m), C.int(steps), C.int64_t(nFilters), C.int(device), ptr(R), ptr(H), shared,
		ptr(realObs), ptr(computedObs), C.GKB_HOST, ptr(xHat0), ptr(P0), nil)
	return xHat0, P0, lastErr(rc)
}

// HouseholderTransf mirrors gokalman.HouseholderTransf (helper.go:142-172), in place on A ((n+m) x (n+1)).
func HouseholderTransf(A *mat64.Dense, n, m int) error {
	data := raw(A)
	if err := lastErr(C.gkb_householder_transf(C.int(n), C.int(m), 1, 0, ptr(data), C.GKB_HOST)); err != nil {
		return err
	}
	r, c := A.Dims()
	for i := 0; i < r; i++ {
		for j := 0; j < c; j++ {
			A.Set(i, j, data[i*c+j])
		}
	}
	return nil
}

// FilterMajor reports the array layout of a handle: large-state Vanilla handles (n = 16, 24, ... 64) take and
// return filter-major arrays [N][C] instead of [C][N] (see gkb_filter_major in the header).
func (kf *Vanilla) FilterMajor() bool { return C.gkb_filter_major(kf.h) != 0 }

// ---- round 2 entry points (same caveat: source only, never compiled here) --------------------------------

// TestedModel is the tested filter's OWN model for ChiSquareWith (chisquare.go:16 takes any LDKF: its F / G / H and
// its Noise's Q / R); nil slices mean "the truth generator's matrix".
type TestedModel struct{ F, G, H, Q, R []float64 }

// ChiSquareWith is ChiSquare with a tested filter that carries its own model, on one or several GPUs of this
// process.  devices == nil: one GPU (cfg.device); otherwise the trials are sharded over `devices` and the per-step
// sums are reduced inside the C-ABI (peer = the rank-ordered NVLink peer-memory sum, else one ncclAllReduce).
func ChiSquareWith(cfg *C.gkb_mc_config, tested *TestedModel, devices []int32, peer bool, steps int) (nis, nees []float64, err error) {
	if tested != nil {
		cfg.filter_F, cfg.filter_G, cfg.filter_H, cfg.filter_Q, cfg.filter_R = ptr(tested.F), ptr(tested.G), ptr(tested.H), ptr(tested.Q), ptr(tested.R)
	}
	nis, nees = make([]float64, steps), make([]float64, steps)
	var first C.int32_t
	out := C.gkb_mc_outputs{mem: C.GKB_HOST, nis: ptr(nis), nees: ptr(nees), first_error: &first}
	var rc C.int
	if devices == nil {
		rc = C.gkb_mc_chisquare(cfg, &out)
	} else {
		mode := C.int(C.GKB_REDUCE_NCCL)
		if peer {
			mode = C.GKB_REDUCE_PEER
		}
		rc = C.gkb_mc_chisquare_multi(cfg, (*C.int)(unsafe.Pointer(&devices[0])), C.int(len(devices)), mode, &out)
	}
	if rc != 0 {
		return nil, nil, lastErr(rc)
	}
	if first != 0 {
		panic("a trial's Update failed") // chisquare.go:40-42 panics
	}
	return nis, nees, nil
}

// SetAWGN arms the handle with AWGN noise (noise.go:109-159) drawn on the device: Process(k) / Measurement(k) of
// filter f at step k are Philox(seed, filterOffset + f, k) normals coloured with chol(Q), chol(R).
func (kf *Vanilla) SetAWGN(Q, R mat64.Symmetric, seed uint64, filterOffset int64) error {
	mr, _ := R.Dims()
	if rc := C.gkb_set_noise(kf.h, ptr(raw(Q)), C.int(mr), ptr(raw(R))); rc != 0 {
		return lastErr(rc)
	}
	return lastErr(C.gkb_set_philox_noise(kf.h, C.uint64_t(seed), C.int64_t(filterOffset)))
}

// VanLoanBatch is c2d.go:13-75 for `count` systems (A [n*n][count] or shared, dt [count] or one value).
func VanLoanBatch(n, q int, count int64, device int, A []float64, aShared bool, Gamma []float64, gShared bool, W, dt []float64,
	dtShared bool) (F, Q []float64, err error) {
	F, Q = make([]float64, int64(n*n)*count), make([]float64, int64(n*n)*count)
	b := func(v bool) C.int {
		if v {
			return 1
		}
		return 0
	}
	rc := C.gkb_van_loan(C.int(n), C.int(q), C.int64_t(count), C.int(device), ptr(A), b(aShared), ptr(Gamma), b(gShared), ptr(W),
		ptr(dt), b(dtShared), C.GKB_HOST, ptr(F), ptr(Q), nil)
	return F, Q, lastErr(rc)
}

// HybridHandle is the NLDKF handle of NewHybridKF (gkb_create_hybrid).
type HybridHandle struct{ h *C.gkb_filter }

// ---- the NLDKF interface (kalman.go:46-62) on the GPU: the headline path ------------------------------------------
//
// HybridKF satisfies gokalman.NLDKF: Prepare / PreparePNT stash the epoch's matrices exactly like hybrid.go:78-89,
// Update / Predict make ONE gkb_nl_run(steps = 1) call with them (flags = GKB_F_MEAS | GKB_F_EKF | GKB_F_SNC as the
// reference's state says) and lock the filter again (hybrid.go:140,201-203).  With nFilters = 1 it is a drop-in for the
// Go object; RunBatch below is the batched form the benchmark measures (N filters x steps epochs in one call).
type HybridKF struct {
	HybridHandle
	n, m, q  int
	noise    gokalman.Noise
	phi, ht  []float64 // this epoch's Phi (n x n) and Htilde (m x n), row-major
	gamma    []float64 // PreparePNT's Gamma (n x q) or nil
	locked   bool
	ekf, snc bool
}

// NewHybridKF mirrors hybrid.go:23-34.  nFilters = 1 is the drop-in filter: the handle starts in reference-order
// arithmetic (gkb_create_hybrid), so every Estimate equals the reference's bit for bit; a batch (nFilters > 1) starts in
// the fast FMA mode -- SetStrict(true) selects the reference's arithmetic for it (what ill-conditioned OD runs need).
func NewHybridKF(x0 *mat64.Vector, P0 mat64.Symmetric, noise gokalman.Noise, measSize int, nFilters, device int) (*HybridKF, error) {
	n, _ := P0.Dims()
	q, _ := noise.ProcessMatrix().Dims()
	kf := &HybridKF{n: n, m: measSize, q: q, noise: noise, locked: true}
	rc := C.gkb_create_hybrid(C.int(n), C.int(measSize), C.int(q), C.int64_t(nFilters), C.int(device), ptr(raw(x0)), 0,
		ptr(raw(P0)), ptr(raw(noise.ProcessMatrix())), ptr(raw(noise.MeasurementMatrix())), &kf.h)
	return kf, lastErr(rc)
}

func (kf *HybridKF) EKFEnabled() bool { return kf.ekf }
func (kf *HybridKF) EnableEKF()       { kf.ekf = true }
func (kf *HybridKF) DisableEKF()      { kf.ekf = false }
func (kf *HybridKF) SetNoise(n gokalman.Noise) {
	kf.noise = n
	R := n.MeasurementMatrix()
	mr, _ := R.Dims()
	C.gkb_set_noise(kf.h, ptr(raw(n.ProcessMatrix())), C.int(mr), ptr(raw(R)))
}
func (kf *HybridKF) Prepare(Phi, Htilde *mat64.Dense) { // hybrid.go:78-82
	kf.phi, kf.ht, kf.locked = raw(Phi), raw(Htilde), false
}
func (kf *HybridKF) PreparePNT(Gamma *mat64.Dense) { // hybrid.go:86-89
	kf.gamma, kf.snc = raw(Gamma), true
}
func (kf *HybridKF) Update(real, computed *mat64.Vector) (gokalman.Estimate, error) {
	return kf.step(C.GKB_F_MEAS, raw(real), raw(computed))
}
func (kf *HybridKF) Predict() (gokalman.Estimate, error) { return kf.step(0, nil, nil) }

func (kf *HybridKF) step(fl C.int, real, computed []float64) (gokalman.Estimate, error) {
	if kf.locked { // hybrid.go:105-107
		return nil, fmt.Errorf("kf is locked (call Prepare first)")
	}
	if kf.ekf {
		fl |= C.GKB_F_EKF
	}
	if kf.snc {
		fl |= C.GKB_F_SNC
	}
	flags := []C.uint8_t{C.uint8_t(fl)}
	est := &Estimate{n: kf.n, m: kf.m, state: make([]float64, kf.n), meas: make([]float64, kf.m), innov: make([]float64, kf.m),
		covar: make([]float64, kf.n*kf.n), pred: make([]float64, kf.n*kf.n), gain: make([]float64, kf.n*kf.m)}
	var status C.int32_t
	out := C.gkb_outputs{mem: C.GKB_HOST, state: ptr(est.state), meas: ptr(est.meas), innov: ptr(est.innov),
		covar: ptr(est.covar), pred_covar: ptr(est.pred), gain: ptr(est.gain), status: &status}
	// one filter: its Phi / Htilde / Gamma are "shared by the batch" (phi_shared = h_shared = 1)
	rc := C.gkb_nl_run(kf.h, 1, &flags[0], ptr(kf.phi), 1, ptr(kf.ht), 1, ptr(real), ptr(computed), ptr(kf.gamma), C.GKB_HOST, &out)
	kf.locked, kf.snc = true, false // hybrid.go:140, 201-203
	if rc != 0 {
		return nil, lastErr(rc)
	}
	if status != 0 { // the reference returns (nil, err): singular S, asymmetric covariance (strict mode)
		return nil, fmt.Errorf("gokalman_b200: update failed with status %d", int(status))
	}
	return est, nil
}

// RunBatch advances every filter of the handle through `steps` epochs in one call: per-filter streams laid out
// [steps][component][nFilters] (Phi 36, Htilde 12, real 2, computed 2 at n = 6, m = 2), flags[k] per epoch shared by the
// batch.  This is the call bench.py times (device-resident streams: in_mem / out.mem = GKB_DEVICE).
func (kf *HybridKF) RunBatch(steps int, flags []uint8, Phi, Htilde, real, computed, Gamma []float64, inMem C.int, out *C.gkb_outputs) error {
	return lastErr(C.gkb_nl_run(kf.h, C.int(steps), (*C.uint8_t)(unsafe.Pointer(&flags[0])), ptr(Phi), 0, ptr(Htilde), 0,
		ptr(real), ptr(computed), ptr(Gamma), inMem, out))
}

// SRIF (srif.go:14-49) is the same binding over gkb_create_srif; EnableEKF / DisableEKF / PreparePNT are the reference's no-ops.
func NewSRIF(x0 *mat64.Vector, P0 mat64.Symmetric, measSize int, nonTriR bool, n gokalman.Noise, nFilters, device int) (*HybridKF, error) {
	dim, _ := P0.Dims()
	kf := &HybridKF{n: dim, m: measSize, noise: n, locked: true}
	tri := C.int(0)
	if nonTriR {
		tri = 1
	}
	rc := C.gkb_create_srif(C.int(dim), C.int(measSize), C.int64_t(nFilters), C.int(device), ptr(raw(x0)), 0, ptr(raw(P0)),
		ptr(raw(n.MeasurementMatrix())), tri, &kf.h)
	return kf, lastErr(rc)
}

// SetStrict selects the reference-order arithmetic (bit-identical to the CPU restatement of hybrid.go:104-204).
func (kf *HybridHandle) SetStrict(on bool) error {
	v := C.int(0)
	if on {
		v = 1
	}
	return lastErr(C.gkb_set_strict(kf.h, v))
}

// RunOD is the fused OD run: the step the reference's callers do with the smd propagator before every
// Prepare(Phi, Htilde) + Update(real, computed) (hybrid_test.go:159-294) happens on the device, per filter.
func (kf *HybridHandle) RunOD(cfg *C.gkb_od_config, steps int, flags []uint8, out *C.gkb_outputs) error {
	return lastErr(C.gkb_od_run(kf.h, cfg, C.int(steps), (*C.uint8_t)(unsafe.Pointer(&flags[0])), out))
}
