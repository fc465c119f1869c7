#!/bin/bash
# round 2, GPU call Z: the 4-CTA-cluster experiment kept behind CZ_CLUSTER4: full GPU tier (default path re-validated) + smoke
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 1200 python -m pytest tests -m gpu -x -q -s) > $GOUT/z_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/z_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $GOUT/z_smoke.log 2>&1
ls -la $GOUT
