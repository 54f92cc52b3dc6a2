run() { # label, env...
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label',d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
}
run cga1_res0 NCCL_CGA_CLUSTER_SIZE=1 DGRM_RESERVE_SMS=0
run cga1_res4 NCCL_CGA_CLUSTER_SIZE=1 DGRM_RESERVE_SMS=4
run cga1_nt128_res0 NCCL_CGA_CLUSTER_SIZE=1 NCCL_NTHREADS=128 DGRM_RESERVE_SMS=0
run cga1_res8 NCCL_CGA_CLUSTER_SIZE=1 DGRM_RESERVE_SMS=8
