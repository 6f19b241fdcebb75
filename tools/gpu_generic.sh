#!/bin/bash
# generic path and bodies: numbers for profiles/ (one GPU)
set -u
mkdir -p gpurun_out
: > gpurun_out/r02h_generic.jsonl
timeout 300 python tools/generic_bench.py --dim 2 --size 1024 --steps 10 >> gpurun_out/r02h_generic.jsonl 2>gpurun_out/r02h_generic.err
MMF_GENERIC_FUSED=0 timeout 300 python tools/generic_bench.py --dim 2 --size 1024 --steps 10 >> gpurun_out/r02h_generic.jsonl 2>>gpurun_out/r02h_generic.err
timeout 300 python tools/generic_bench.py --two-level 48 --steps 10 >> gpurun_out/r02h_generic.jsonl 2>>gpurun_out/r02h_generic.err
timeout 300 python tools/generic_bench.py --size 128 --steps 10 >> gpurun_out/r02h_generic.jsonl 2>>gpurun_out/r02h_generic.err
timeout 600 python tools/generic_bench.py --size 192 --steps 10 --bodies >> gpurun_out/r02h_generic.jsonl 2>>gpurun_out/r02h_generic.err
cat gpurun_out/r02h_generic.jsonl; tail -3 gpurun_out/r02h_generic.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_ -s 12 -c 4 -f -o gpurun_out/r02h_generic python tools/generic_bench.py --dim 2 --size 1024 --steps 2 > gpurun_out/r02h_generic_ncu.log 2>&1
echo "ncu exit $?"
