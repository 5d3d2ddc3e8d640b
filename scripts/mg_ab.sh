#!/bin/bash
# A/B of exchange variants on N GPUs (scripts/mg_diag.py under each library); stops at the first variant that fails or hangs.
N=${N:-2}
for v in "$@"; do
  if [ $v = main ]; then L=tfdiffeq_b200/libb2ode.so; else L=tfdiffeq_b200/variants/libb2ode_$v.so; fi
  echo "== $v"
  B2ODE_LIB=$L timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/mg_diag.py > /tmp/mg_$v.log 2>&1
  rc=$?
  grep "rank 0" /tmp/mg_$v.log | cut -c1-140
  if [ $rc != 0 ]; then echo "variant $v FAILED rc=$rc"; tail -5 /tmp/mg_$v.log; exit 1; fi
done
