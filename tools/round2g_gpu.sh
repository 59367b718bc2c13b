#!/bin/bash
# Round 2, the very last GPU pass (about 60 s of box time left): the column variants of one motif on views.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
( time timeout 30 python -m pytest tests/test_gpu_refine.py -x -q -p no:cacheprovider ) > $O/r02g_pytest_refine.log 2>&1
echo "refine (one context, refactored tail) rc=$?" | tee -a $O/r02g_status.txt
( time timeout 30 python -m pytest tests/test_gpu_motif_batch.py -x -q -p no:cacheprovider -k variants ) > $O/r02g_pytest_variants.log 2>&1
echo "variants on views rc=$?" | tee -a $O/r02g_status.txt
( time timeout 30 python tools/variant_batch.py --sizes 100000,1000000 ) > $O/r02g_variant_batch.json 2> $O/r02g_variant_batch.err
echo "variant_batch rc=$?" | tee -a $O/r02g_status.txt
cat $O/r02g_variant_batch.json
tail -n 3 $O/r02g_pytest_refine.log
tail -n 3 $O/r02g_pytest_variants.log
