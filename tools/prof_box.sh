#!/bin/bash
# Runs ON the GPU box (via gpurun): one ncu --set full pass over the first frames of bench.py, condensed to text there
# (the .ncu-rep of several kernels exceeds what gpurun copies back).  usage: tools/prof_box.sh <tag> [launch-count]
tag=${1:-prof}; cnt=${2:-6}; kreg=${KREG:-"me_picture|tq_batch|mc_batch|sao_stats|ois_picture|ois_dc|sao_apply|dlf_pass"}; skip=${SKIP:-16}
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on \
  --kernel-name regex:"$kreg" --launch-skip $skip --launch-count $cnt \
  -o /tmp/$tag -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.log 2>&1
python tools/ncu_summary.py /tmp/$tag.ncu-rep > gpurun_out/${tag}_summary.txt 2>&1
python tools/ncu_lines.py /tmp/$tag.ncu-rep "me_picture:k12817me_picture_kernelILb0" me.cu 70 > gpurun_out/${tag}_lines_me0.txt 2>&1
python tools/ncu_lines.py /tmp/$tag.ncu-rep "me_picture:k25617me_picture_kernelILb1" me.cu 50 > gpurun_out/${tag}_lines_me1.txt 2>&1
python tools/ncu_lines.py /tmp/$tag.ncu-rep "tq_batch:tq_batch_kernelILb0" tq.cu 60 > gpurun_out/${tag}_lines_tq.txt 2>&1
python tools/ncu_lines.py /tmp/$tag.ncu-rep "mc_batch8:mc_batch8_kernelILb1" mc.cu 400 > gpurun_out/${tag}_lines_mc.txt 2>&1
python tools/ncu_lines.py /tmp/$tag.ncu-rep "sao_stats2_kernel" sao.cu 40 > gpurun_out/${tag}_lines_sao.txt 2>&1
python tools/ncu_lines.py /tmp/$tag.ncu-rep "ois_dc_kernel" intra.cu 30 > gpurun_out/${tag}_lines_ois.txt 2>&1
python tools/ncu_traffic.py /tmp/$tag.ncu-rep gpurun_out/${tag}_traffic.json > gpurun_out/${tag}_traffic.log 2>&1
ls -la /tmp/$tag.ncu-rep
sz=$(stat -c %s /tmp/$tag.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 40000000 ]; then cp /tmp/$tag.ncu-rep gpurun_out/; fi
