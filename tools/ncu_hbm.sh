#!/bin/bash
# ncu --set full captures of the HBM / gather kernels of one training step (2 launches each).  The .ncu-rep files stay on
# the GPU box (gpurun_out is capped at 64 MiB): only the raw-page CSV of each capture comes back, as
# gpurun_out/ncu/<kernel>[--bf16].csv, for tools/ncu_hbm_table.py.   usage (under gpurun): tools/ncu_hbm.sh [--bf16]
mkdir -p gpurun_out/ncu
KERNELS=${KERNELS:-"deform_psroi_fwd_sep deform_psroi_bwd deform_im2col mpt_decode mpt_nms_fast mpt_nms_assign colsum_kernel bn_relu_bwd_apply affine_act stem_conv sgd_mom_dev maxpool"}
for k in $KERNELS; do
  ncu --set full --clock-control none --profile-from-start off -k regex:$k -c 2 -f -o /tmp/ncu_$k \
      python tools/profile_step.py $1 > /dev/null 2>&1
  if [ -f /tmp/ncu_$k.ncu-rep ]; then
    ncu -i /tmp/ncu_$k.ncu-rep --page raw --csv > gpurun_out/ncu/$k$1.csv 2>/dev/null
    rm -f /tmp/ncu_$k.ncu-rep
  fi
done
ls -la gpurun_out/ncu | tail -30
