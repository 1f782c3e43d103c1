#!/bin/bash
# registers / spills per kernel of one .cu file:  scripts/ptxas_report.sh mbpls_b200/csrc/fused.cu
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -c "$1" -o /tmp/ptxas_report.o 2>&1 | python3 -c "
import sys,re,subprocess
cur=None; sp=''
for ln in sys.stdin.read().splitlines():
    m=re.search(r\"Compiling entry function '(\S+)'\",ln)
    if m:
        cur=subprocess.run(['c++filt',m.group(1)],capture_output=True,text=True).stdout.strip()
        cur=re.sub(r'\(anonymous namespace\)::','',cur); cur=re.sub(r'\((?:\(anonymous namespace\)::)?\w+Args\)','',cur)
    elif 'spill' in ln: sp=ln.strip()
    elif 'Used' in ln and cur: print(cur,'|',ln.split(':',1)[1].strip(),'|',sp); cur=None
    elif 'error' in ln: print(ln)
"
