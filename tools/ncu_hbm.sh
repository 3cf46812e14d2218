#!/bin/bash
# ncu --set full captures of the HBM / gather kernels of one training step (2 launches each) -> gpurun_out/ncu_<name>.ncu-rep
# usage (under gpurun): tools/ncu_hbm.sh [--bf16]
mkdir -p gpurun_out
for k in deform_psroi_fwd_sep deform_psroi_bwd deform_im2col mpt_decode mpt_nms_fast mpt_nms_assign colsum_kernel bn_relu_bwd_apply affine_act stem_conv sgd_mom_dev cast_rows; do
  ncu --set full --clock-control none --profile-from-start off -k regex:$k -c 2 -f -o gpurun_out/ncu_$k$1 \
      python tools/profile_step.py $1 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
