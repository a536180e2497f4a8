#!/bin/bash
# Compile the library with ptxas -v and print registers / spills / shared memory per kernel matching $1.
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas=-v -I include \
  -o ocpg_b200/lib/libmsda_sm100.so ocpg_b200/csrc/msda_sm100.cu 2>&1 | grep -v "^$" > /tmp/ptxas.log
grep -E "error|warning" /tmp/ptxas.log | head -20
python - "$1" <<'PY'
import re, sys, subprocess
pat = sys.argv[1] if len(sys.argv) > 1 else "msda"
txt = open("/tmp/ptxas.log").read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
    if pat in name:
        print(f"{name:70s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill {m.group(3)}/{m.group(4)}")
PY
