#!/bin/bash
# Round 2, last GPU pass (about 90 s of box time left): the DevBuf / view changes on the device.
# Order = importance; every step writes its own log so that a cut-off call still leaves what ran.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
( time timeout 40 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider ) > $O/r02f_pytest_parity.log 2>&1
echo "parity rc=$?" | tee -a $O/r02f_status.txt
( time timeout 45 python -m pytest tests/test_gpu_motif_batch.py -x -q -p no:cacheprovider ) > $O/r02f_pytest_views.log 2>&1
echo "views rc=$?" | tee -a $O/r02f_status.txt
( time timeout 40 python tools/motif_batch.py --shapes cfg1,cfg2 --contexts 4 --motifs 16 ) > $O/r02f_motif_batch.json 2> $O/r02f_motif_batch.err
echo "motif_batch rc=$?" | tee -a $O/r02f_status.txt
tail -3 $O/r02f_pytest_parity.log $O/r02f_pytest_views.log
cat $O/r02f_motif_batch.json
( time timeout 25 python tools/motif_batch.py --shapes cfg2 --contexts 8 --motifs 16 ) > $O/r02f_motif_batch_8ctx.json 2> $O/r02f_motif_batch_8ctx.err
echo "motif_batch8 rc=$?" | tee -a $O/r02f_status.txt
cat $O/r02f_motif_batch_8ctx.json
