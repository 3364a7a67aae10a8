#!/bin/bash
# usage: tools/ptxas_info.sh <file.cu> [filter]   -- registers / spills / smem per kernel instantiation
f=$1; pat=${2:-.}
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --use_fast_math \
  -I include -I mamba_tts_project_b200/csrc -Xptxas -v -c "$f" -o /tmp/_ptxas_info.o 2>&1 | python3 -c '
import sys,re,subprocess
name=None; spill=""
for line in sys.stdin:
    m=re.search(r"Compiling entry function .(\S+). for",line)
    if m:
        name=subprocess.run(["c++filt",m.group(1)],capture_output=True,text=True).stdout.strip()
        name=re.sub(r"\(.*","",name); continue
    m=re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads",line)
    if m: spill="stack=%s spill=%s/%s"%m.groups(); continue
    m=re.search(r"Used (\d+) registers(.*)",line)
    if m and name:
        print(name, "regs=%s"%m.group(1), spill, m.group(2).strip()); name=None
    elif "error" in line or "warning" in line: print(line.rstrip())
' | grep -E "$pat"
