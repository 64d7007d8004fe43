# usage: bash scripts/scale_run.sh N   (the bench lines of one box with N GPUs: weak scaling of the three workloads + SIFT strong scaling)
N=$1
for w in sift bow orb; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --workload $w --gpus $N --steps 5 --warmup 3 > gpurun_out/r2z_bench_${w}_n$N.json 2> gpurun_out/r2z_${w}_n$N.err
done
[ "$N" = 8 ] || python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --scaling strong --steps 5 --warmup 3 > gpurun_out/r2z_bench_sift_strong_n$N.json 2> gpurun_out/r2z_sift_strong_n$N.err
python scripts/show_bench.py gpurun_out/r2z_bench_*_n$N.json
