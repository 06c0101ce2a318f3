#!/bin/bash
# usage: ab.sh name1=lib1 name2=lib2 ... ; prints step, entropy, granule ms (bit-exact) per library variant ("default" = product library)
# CAUTION: the product library is rebuilt on import whenever a source is newer than it -- on the GPU box "default" is therefore the
# WORKING TREE, not the last commit.  Compare explicit variants (tools/variant.py base, built before editing) unless the tree is clean.
for kv in "$@"; do
  name=${kv%%=*}; lib=${kv#*=}
  if [ "$lib" = "default" ]; then unset L3B_LIB; else export L3B_LIB=$PWD/$lib; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fused --parity-sample 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', 'step', round(d['ms_per_step'],2), 'entropy', round(d['roofline']['entropy_kernels_ms'],2), 'granule', round(d['roofline']['ms_per_launch'],2), 'parity', d['parity']['oracle_sample_mismatches'])"
done
