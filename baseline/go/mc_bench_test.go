package gokalmanbench

// Goroutine-sharded Monte Carlo + chi-square over the REAL gokalman package: the CPU baseline BASELINE.json asks
// for ("the reference's goroutine-parallel Monte Carlo ... with GOMAXPROCS and core count stated").  The reference's
// own harness is a serial double loop over one stateful filter (montecarlo.go:108-117), so the sharding is done
// here: every goroutine owns its own pure predictor + tested filter and runs trials/GOMAXPROCS trials.
//
// SOURCE ONLY (no Go toolchain in the build image).  Run on a machine with Go:
//     go test -bench MonteCarloJerk3 -benchtime 1x -cpu 1,8,16
// and compare "updates/s" with bench.py's cpu_baseline (the C restatement) and `value` (the GPU).

import (
	"runtime"
	"sync"
	"testing"

	"github.com/ChristopherRabotin/gokalman"
	"github.com/gonum/matrix/mat64"
)

// jerk3 is the fixture of montecarlo_test.go:12-26 / helper_test.go:17-22.
func jerk3() (F, G, H *mat64.Dense, Q, R *mat64.SymDense, x0 *mat64.Vector, P0 *mat64.SymDense) {
	F = mat64.NewDense(3, 3, []float64{1, 0.01, 5e-5, 0, 1, 0.01, 0, 0, 1})
	G = mat64.NewDense(3, 1, []float64{5e-7 / 3, 5e-5, 0.01})
	H = mat64.NewDense(1, 3, []float64{1, 0, 0})
	Q = mat64.NewSymDense(3, []float64{2.5e-15, 6.25e-13, 25e-11 / 3, 6.25e-13, 5e-7 / 3, 2.5e-8, 25e-11 / 3, 2.5e-8, 5e-6})
	R = mat64.NewSymDense(1, []float64{0.5})
	x0 = mat64.NewVector(3, []float64{0, 0.35, 0})
	P0 = gokalman.ScaledIdentity(3, 10)
	return
}

func BenchmarkMonteCarloJerk3(b *testing.B) {
	const trials, steps = 10000, 1000
	workers := runtime.GOMAXPROCS(0)
	F, G, H, Q, R, x0, P0 := jerk3()
	controls := []*mat64.Vector{mat64.NewVector(1, nil)} // a single vector = zero controls (montecarlo.go:98-104)
	b.ResetTimer()
	for it := 0; it < b.N; it++ {
		var wg sync.WaitGroup
		for w := 0; w < workers; w++ {
			wg.Add(1)
			go func() {
				defer wg.Done()
				mckf, _, _ := gokalman.NewPurePredictorVanilla(x0, P0, F, G, H, gokalman.NewAWGN(Q, R))
				runs := gokalman.NewMonteCarloRuns(trials/workers, steps, 1, controls, mckf)
				kf, _, _ := gokalman.NewVanilla(x0, P0, F, G, H, gokalman.NewNoiseless(Q, R))
				if _, _, err := gokalman.NewChiSquare(kf, runs, controls, true, true); err != nil {
					b.Error(err)
				}
			}()
		}
		wg.Wait()
	}
	b.ReportMetric(float64(trials/workers*workers*steps*b.N)/b.Elapsed().Seconds(), "updates/s")
	b.ReportMetric(float64(workers), "GOMAXPROCS")
	b.ReportMetric(float64(runtime.NumCPU()), "cores")
}
