#!/usr/bin/env bash
# BASELINE configs[0] through the reference's OWN programs and their own gettimeofday timers: deo_doe_test and
# inverter_multishift_test (benchmark mode: 15 equal shifts, MaxCGIterations iterations) at 8^4, once as the pure-reference
# gcc CPU build and once linked against libstaple_b200.so (oracle/build_ref_host.sh), on the same box.
# usage: scripts/gpu_ref_host_timing.sh [N0xN1xN2xN3 (default 8x8x8x8; the binaries must exist: oracle/build_ref_host.sh)] [iterations]
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
GEOM=${1:-8x8x8x8}; ITERS=${2:-200}
IFS=x read -r NX NY NZ NT <<< "$GEOM"
OUT=$ROOT/gpurun_out/ref_host_timing_$GEOM.txt
mkdir -p "$ROOT/gpurun_out"; : > "$OUT"
export LD_LIBRARY_PATH=${LD_LIBRARY_PATH:-}:/usr/local/cuda/lib64
for kind in staple ref; do
  T=$(mktemp -d); cd "$T"
  sed -e "s/^\(DeoDoeIterations *\)[0-9]*/\1$ITERS/" -e 's/^\(SaveResults *\)[0-9]*/\10/' -e "s/^\(MaxCGIterations *\)[0-9]*/\1$ITERS/" \
      -e 's/^\(MultiShiftInverterRepetitions *\)[0-9]*/\13/' -e "s/^nx .*/nx $NX/" -e "s/^ny .*/ny $NY/" -e "s/^nz .*/nz $NZ/" -e "s/^nt .*/nt $NT/" \
      "$ROOT/tests/golden/ref_host/deo_doe_8x8x8x8.set" > in.set
  python - "$ROOT" <<'PY'
import json, sys, os
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from test_gpu_reference_host import _remez_text
for name, r in json.load(open(os.path.join(sys.argv[1], "tests/golden/ref_host/ratapproxes.json"))).items():
    open(name, "w").write(_remez_text(r))
PY
  for prog in deo_doe_test inverter_multishift_test; do
    echo "=== $prog ($kind) ===" >> "$OUT"
    [ -x "$ROOT/oracle/_ref/${prog}_${kind}_$GEOM" ] || { echo "missing binary" >> "$OUT"; continue; }
    timeout 1200 "$ROOT/oracle/_ref/${prog}_${kind}_$GEOM" in.set 2> err.log | grep -E "PRECISION|Time for 1|hot path" >> "$OUT"
    grep "hot path" err.log >> "$OUT"
  done
  cd /; rm -rf "$T"
done
cat "$OUT"
