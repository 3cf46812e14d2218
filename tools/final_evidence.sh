#!/bin/bash
# Round-end evidence run on one B200 (under gpurun): smoke, the GPU suite, the default bench line with the per-shape and
# per-entry-point tables, ncu launch lists of one TF32 and one bf16 step, ncu --set full summaries of three tcgen05 shapes
# and of the BatchNorm kernels.  Everything lands in gpurun_out/ as text (the .ncu-rep files stay on the box).
mkdir -p gpurun_out/ncu
python __graft_entry__.py --smoke 2>&1 | tail -1
python -m pytest tests -m gpu -q 2>&1 | tail -2
SNIPER_DUMP_GEMM=gpurun_out/gemm_shapes_r02_v2.md SNIPER_BREAKDOWN=gpurun_out/entry_points_r02_v2.md \
  python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_err.log
echo "bench rc=$?"; tail -2 gpurun_out/bench_err.log
for m in "" "--bf16"; do
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file /tmp/launches$m.csv python tools/profile_step.py $m > /dev/null 2>&1
  python tools/summarize_launches.py /tmp/launches$m.csv > gpurun_out/launches_r02_v2$m.md
done
for shp in "20480 256 2304 stats" "20480 1024 256 stats" "20480 1024 256 bf16 stats" "20480 3072 4608"; do
  tag=$(echo $shp | tr ' ' '_')
  ncu --set full --clock-control none --import-source on --profile-from-start off -c 1 -f -o /tmp/ncu_g_$tag \
      python tools/gemm_one.py $shp > /dev/null 2>&1
  ncu -i /tmp/ncu_g_$tag.ncu-rep --page raw --csv > gpurun_out/ncu/gemm_$tag.csv 2>/dev/null
  ncu -i /tmp/ncu_g_$tag.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Tensor|Issue Slots|Registers|Dynamic Shared|L2 Hit|DRAM|Executed Ipc|Warp Cycles Per Issued" > gpurun_out/ncu/gemm_$tag.txt
  rm -f /tmp/ncu_g_$tag.ncu-rep
done
KERNELS="colsum_kernel bn_relu_bwd_apply bn_apply_train stem_im2col" tools/ncu_hbm.sh > /dev/null 2>&1
KERNELS="colsum_kernel bn_relu_bwd_apply bn_apply_train" tools/ncu_hbm.sh --bf16 > /dev/null 2>&1
ls gpurun_out/ncu | head -30
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_final.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_iterator"]["value"], d["config3"]["value"],
      d["roofline"]["frac"], d["config3"]["roofline"]["frac"], d["cpu_baseline"]["value"], d["gpu_launches"])
PY
