#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
OKL_NVLS=1 run 29521 --steps 20 --no-configs > gpurun_out/dpq_cifar_nvls.json 2> gpurun_out/dpq_cifar_nvls.err; echo "cifar nvls exit $?"; grep -h "okl\]" gpurun_out/dpq_cifar_nvls.err gpurun_out/dpq_cifar_nvls.json | head -5
OKL_NVLS=0 run 29522 --steps 20 --no-configs > gpurun_out/dpq_cifar_p2p.json 2> gpurun_out/dpq_cifar_p2p.err; echo "cifar p2p exit $?"
OKL_NVLS=1 run 29523 --workload mnist_cnn --precision tf32 --steps 500 --no-configs > gpurun_out/dpq_mnist_nvls.json 2> gpurun_out/dpq_mnist_nvls.err; echo "mnist nvls exit $?"
OKL_NVLS=0 run 29524 --workload mnist_cnn --precision tf32 --steps 500 --no-configs > gpurun_out/dpq_mnist_p2p.json 2> gpurun_out/dpq_mnist_p2p.err; echo "mnist p2p exit $?"
python - <<'PY'
import json
for f in ("cifar_nvls","cifar_p2p","mnist_nvls","mnist_p2p"):
    try:
        txt=[l for l in open(f"gpurun_out/dpq_{f}.json").read().strip().splitlines() if l.startswith("{")]
        l=json.loads(txt[-1])
        print(f, {k:l.get(k) for k in ("value","ms_per_step","n_gpus","scaling","other_scaling")}, "dp", l.get("dp_check") and {k:l["dp_check"][k] for k in ("param_max_abs_diff_across_ranks","max_rel")})
    except Exception as e:
        print(f, "parse fail", e); print(open(f"gpurun_out/dpq_{f}.err").read()[-1500:])
PY
