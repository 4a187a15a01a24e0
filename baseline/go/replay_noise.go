// Package gokalmanbench holds reference-side helpers for validating and timing the B200 engine against the REAL
// gokalman package.  SOURCE ONLY: the build image has no Go toolchain (and no gonum), so nothing here is ever
// compiled or run by this repository; it is for a maintainer with Go installed.
package gokalmanbench

import (
	"github.com/gonum/matrix/mat64"
)

// ReplayNoise implements gokalman.Noise (noise.go:13-20) over pre-baked samples indexed by step, keeping Q and R
// (gokalman.BatchNoise reports Q = R = 0, noise.go:88-99, and so cannot drive a real filter).  Feed it the
// arrays the engine dumps (MonteCarloRuns.Truth(with_noise=True): w [steps][n][trials], v [steps][m][trials]) to run
// the reference filters on exactly the samples the GPU used.
type ReplayNoise struct {
	Q, R mat64.Symmetric
	W, V []*mat64.Vector // one vector per step
}

func (n ReplayNoise) Process(k int) *mat64.Vector        { return n.W[k] }
func (n ReplayNoise) Measurement(k int) *mat64.Vector    { return n.V[k] }
func (n ReplayNoise) ProcessMatrix() mat64.Symmetric     { return n.Q }
func (n ReplayNoise) MeasurementMatrix() mat64.Symmetric { return n.R }
func (n ReplayNoise) Reset()                             {}
func (n ReplayNoise) String() string                     { return "ReplayNoise" }
