mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_io.py tests/test_graft.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/t_io.log 2>&1
cat gpurun_out/t_io.log | tail -12
for gc in fp32 bf16; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/prof_ddp_step.py --grad-comm $gc > gpurun_out/ddp2_$gc.txt 2> gpurun_out/ddp2_$gc.err
  grep -v Warning gpurun_out/ddp2_$gc.txt | head -20
done
NCCL_MAX_CTAS=8 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/prof_ddp_step.py > gpurun_out/ddp2_fp32_maxctas8.txt 2> gpurun_out/ddp2_maxctas8.err
head -4 gpurun_out/ddp2_fp32_maxctas8.txt; tail -1 gpurun_out/ddp2_fp32_maxctas8.txt
timeout 300 python tools/bench_infer.py --arch edsr_l --scene 2048 --tile 1024 --overlap 64 --guard 16 --reps 1 2>&1 | tail -1
