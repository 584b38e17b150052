#!/bin/bash
# The 8-GPU record of a round (run under `gpurun --gpus 8`): 8-rank NCCL parity tests, the bench with its in-run parity
# self-check, the reference's own train_model on 8 GPUs (BASELINE config 5) at C = 2M and C = 10,575 with the DDP
# gradient check, the collectives timed alone, and the cfg5 bench line.  Outputs under gpurun_out/.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_sharded.py -q -k 'sharded_matches_oracle and 8' > gpurun_out/r2_pytest_8gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_8gpu.log
timeout 150 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; tail -c 1800 gpurun_out/r2_bench_8gpu.json; tail -3 gpurun_out/r2_bench_8gpu.err
timeout 120 $TR --master-port 29522 examples/ref_train_model.py --classes 2000000 --steps 30 --check-grads > gpurun_out/r2_reftrain_8gpu_c2m.log 2>&1; grep -E 'check-grads|ref_train_model:' gpurun_out/r2_reftrain_8gpu_c2m.log
timeout 120 $TR --master-port 29523 examples/ref_train_model.py --classes 10575 --steps 30 > gpurun_out/r2_reftrain_8gpu_c10k.log 2>&1; grep -E 'ref_train_model:' gpurun_out/r2_reftrain_8gpu_c10k.log
timeout 60 $TR --master-port 29524 scripts/time_collectives.py > gpurun_out/r2_collectives_8gpu.json 2>/dev/null; cat gpurun_out/r2_collectives_8gpu.json
timeout 120 $TR --master-port 29525 bench.py --gpus 8 --config cfg1 --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg5_8gpu.json 2> gpurun_out/r2_bench_cfg5_8gpu.err; tail -c 1200 gpurun_out/r2_bench_cfg5_8gpu.json
