#!/bin/bash
# time every K0 v2 A/B build in build_alt/ (tools/build_k0_variants.sh) on one box
mkdir -p gpurun_out
: > gpurun_out/r02_k0_variants.txt
for lib in build_alt/librgbnm_*.so; do
  echo "== $lib" >> gpurun_out/r02_k0_variants.txt
  RGBNM_LIB=$PWD/$lib timeout 300 python tools/k0_prof.py 30 >> gpurun_out/r02_k0_variants.txt 2>/dev/null
done
cat gpurun_out/r02_k0_variants.txt | python -c "
import sys, json
name=None
for ln in sys.stdin:
    if ln.startswith('=='): name=ln.strip()
    elif ln.startswith('{'):
        d=json.loads(ln); print(name, *[(k, d[k]['ms_per_batch_incl_dcstats'], d[k]['frac']) for k in ('eval','train','train_fused_only') if k in d])
"
