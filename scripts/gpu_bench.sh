#!/bin/bash
# bench (both precisions) + ncu launch list + ncu --set full captures of the top kernels.
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --precision fp32 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 exit=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit=$?"
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > gpurun_out/host_cpu.txt
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --cpu-budget 0 --no-graph > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$?"
for k in k_gru k_gemm_bf16_tc k_smpl_verts k_gemm_f32; do
  timeout 600 $NCU --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_$k \
     python bench.py --steps 1 --warmup 3 --cpu-budget 0 --no-graph > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit=$?"
done
head -c 1500 gpurun_out/bench_bf16.json; echo; tail -3 gpurun_out/bench_bf16.err
