#!/bin/bash
# round-2 GPU call M (1 GPU): tests, row-band dependency on / off (slabs, tiles, whole grid)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02m_pytest.log 2>&1
tail -4 gpurun_out/r02m_pytest.log
for rep in 1 2; do
for rd in 0 1; do
OSSG_ROW_DEPS=$rd timeout 120 python tools/slab_time.py 2>&1 | tail -1 | sed "s/^/rowdeps=$rd /"
OSSG_ROW_DEPS=$rd timeout 120 python tools/slab_time.py --xsplit 2>&1 | tail -1 | sed "s/^/rowdeps=$rd /"
OSSG_ROW_DEPS=$rd timeout 120 python tools/variant_time.py 2>&1 | tail -1 | sed "s/^/rowdeps=$rd /"
done
done | tee gpurun_out/r02m_rowdeps.log
