#!/bin/bash
# times K1 (tools/k1_time.py child mode) for every library variant under bftkv_b200/variants/ and the default build
for lib in bftkv_b200/libbftq.so bftkv_b200/variants/*.so; do
  for n in 65536 524288; do
    echo "== $lib n=$n"
    K1_CHILD=1 BFTQ_RSA_KERNEL=r32sq BFTQ_LIB_PATH=$PWD/$lib python tools/k1_time.py $n 2>&1 | tail -1
  done
done
