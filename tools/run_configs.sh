#!/bin/bash
# Runs bench.py on every single-GPU configuration of BASELINE.json (and the dtype / regime variants) and appends the
# JSON lines to $1 (default gpurun_out/configs.jsonl).  One GPU; ~2 minutes.
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/configs.jsonl}
: > "$out"
run() { python bench.py --steps 100 --warmup 10 --no-cpu-baseline "$@" >> "$out" 2>> "${out%.jsonl}.err" || echo "FAILED: $*" >&2; }
run --workload a2d --regime init --dtype f32            # configs[1] (i)
run --workload a2d --regime uniform --dtype f32         #            stress regime
run --workload a2d --regime init --dtype bf16           # configs[1] (ii)
run --workload ytvos --regime init --dtype f32 --input-sets 2 --steps 40   # configs[2] operator
run --workload decoder --regime init --dtype f32        # configs[3]
python tools/bench_encoder.py --shape ytvos --gemm fp32 >> "$out" 2>> "${out%.jsonl}.err"   # configs[2] full encoder, reference GEMM policy
python tools/bench_encoder.py --shape ytvos --gemm tf32 >> "$out" 2>> "${out%.jsonl}.err"
python tools/bench_encoder.py --shape ytvos --gemm tf32 --unfused >> "$out" 2>> "${out%.jsonl}.err"
python tools/bench_encoder.py --shape ytvos --gemm bf16 >> "$out" 2>> "${out%.jsonl}.err"
python tools/bench_encoder.py --shape ytvos --gemm tf32 --dropout 0.1 --graph >> "$out" 2>> "${out%.jsonl}.err"   # training setting, one CUDA graph
python tools/bench_decoder.py --graph >> "$out" 2>> "${out%.jsonl}.err"                                        # configs[3] at layer level
wc -l "$out"
