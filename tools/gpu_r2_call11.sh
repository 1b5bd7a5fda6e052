#!/bin/bash
# Round-2 GPU call 11: PDL edges inside the inference CUDA graph; eager inference with PDL; experimental-switch suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DENSEREG_PDL_GRAPH=1 timeout -s KILL 200 python tools/bench_infer.py --check 2 > gpurun_out/c11_infer_pdlgraph.json 2> gpurun_out/c11_infer.err
DENSEREG_PDL_GRAPH=1 timeout -s KILL 200 python bench.py --config msra_infer --no_cpu_baseline > gpurun_out/c11_bench_infer_pdlgraph.json 2>> gpurun_out/c11_infer.err
timeout -s KILL 200 python tools/bench_infer.py --check 0 --graph 0 > gpurun_out/c11_infer_eager.json 2>> gpurun_out/c11_infer.err
DENSEREG_PDL_GRAPH=1 timeout -s KILL 200 python -m pytest tests/test_gpu_net.py -m gpu -q -k "cuda_graph or infer_end" > gpurun_out/c11_pytest_graph.log 2>&1
DENSEREG_TEST_EXPERIMENTAL=1 timeout -s KILL 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q > gpurun_out/c11_pytest_exp.log 2>&1
cut -c1-900 gpurun_out/c11_infer_pdlgraph.json; cut -c1-400 gpurun_out/c11_bench_infer_pdlgraph.json; cut -c1-600 gpurun_out/c11_infer_eager.json; tail -3 gpurun_out/c11_pytest_graph.log; tail -5 gpurun_out/c11_pytest_exp.log; tail -3 gpurun_out/c11_infer.err
