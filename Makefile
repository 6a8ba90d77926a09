# Convenience targets; everything here is a one-line call into the Python / shell entry points.
PY ?= python

.PHONY: build test test-gpu bench bench-reference smoke sass reference-tests lab clean

build:            ## librlic_b200.so (nvcc, sm_100a) + the CPU checker in oracle/
	$(PY) -c "import __graft_entry__ as g; g.build()"

test:             ## everything that runs without a GPU (a few minutes)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity tests proper, through the public API and the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

smoke:            ## one small invocation of the hot path on cuda:0, checked against the oracle
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench:            ## the contract's JSON line (N = 1; torchrun ... bench.py --gpus N for more)
	$(PY) bench.py --steps 20 --warmup 5

bench-reference:  ## the CPU arm of the same workload
	$(PY) bench.py --impl reference --steps 3 --warmup 1

sass:             ## SASS of the shipped hot kernels and static per-step counts (no GPU)
	sh tools/dump_sass.sh r2
	$(PY) tools/sass_steps.py > profiles/r2_sass_steps.txt

reference-tests:  ## the reference's own, unmodified tests against this package (needs /root/reference)
	$(PY) tools/run_reference_tests.py --backend native
	$(PY) tools/run_reference_tests.py --backend oracle
	$(PY) tools/run_reference_tests.py --backend emulation

lab:              ## the kernel laboratories (tools/kernel_lab, tools/replay_lab)
	sh tools/build_lab.sh

clean:
	rm -f rlic_b200/librlic_b200.so oracle/liblic_oracle.so tests/kernel_emulation/libkernel_emulation.so tools/kernel_lab tools/replay_lab
