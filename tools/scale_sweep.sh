N=$1
for c in hybrid image_vit latent_vit_v2 latent_vit; do
  extra=""; [ "$c" != "hybrid" ] && extra="--no-config5"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --config $c --no-cpu-baseline --no-gpu-eager $extra > gpurun_out/r2_scale${N}_$c.json 2> gpurun_out/r2_scale${N}_$c.err
  tail -c 300 gpurun_out/r2_scale${N}_$c.err | tail -2
done
