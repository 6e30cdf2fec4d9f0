# N-GPU lines of the network workloads and the headline FPS workload (bench.py under torchrun), one JSON line each
N=${1:-8}
mkdir -p gpurun_out/scale
for w in fps fwd fwd_bf16 train; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) \
      bench.py --gpus $N --workload $w --only --no-cpu 2>/dev/null | tail -1 > gpurun_out/scale/n${N}_$w.json
  cut -c1-220 gpurun_out/scale/n${N}_$w.json
done
