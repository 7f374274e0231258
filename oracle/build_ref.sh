#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the UNMODIFIED OpenStaPLE hot-path sources (where they lie under
# /root/reference/src) with gcc into a shared object per compile-time geometry:
#     oracle/_ref/libref_<N0>x<N1>x<N2>x<N3>_r<NRANKS_D3>.so
# following the reference's own gcc recipe (build/compiler_settings_library.txt:47-58:
# gcc -O3 -std=gnu99; OpenACC pragmas ignored => single-threaded CPU code).
# The reference's generator src/double_to_single_transformer.py (Python 2) writes the
# sp_* sources next to its inputs, so a scratch copy of src/ is made under $TMPDIR
# (never inside this repo); only the .so lands in oracle/_ref/ (git-ignored).
# usage: oracle/build_ref.sh N0 N1 N2 N3 [NRANKS_D3=1]
set -euo pipefail
REF=${STAPLE_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
N0=$1; N1=$2; N2=$3; N3=$4; NR=${5:-1}
OUT=$HERE/_ref/libref_${N0}x${N1}x${N2}x${N3}_r${NR}.so
mkdir -p "$HERE/_ref"
[ -d "$REF/src" ] || { echo "reference not present at $REF (prebuilt oracle/_ref is used as is)"; exit 0; }
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/ref_shim.c" ] && [ "$OUT" -nt "$0" ]; then echo "up to date: $OUT"; exit 0; fi
SCR=${STAPLE_ORACLE_SCRATCH:-${TMPDIR:-/tmp}/staple_oracle_src}
if [ ! -f "$SCR/.generated" ]; then
  rm -rf "$SCR"; mkdir -p "$SCR"; cp -r "$REF/src" "$SCR/src"; chmod -R u+w "$SCR"
  ( cd "$SCR/src" \
    && sed -E 's/^(\s*)print (.*)$/\1print(\2)/; s/raw_input/input/' double_to_single_transformer.py > d2s_py3.py \
    && python3 -W ignore d2s_py3.py autoMode silentMode > /dev/null )
  touch "$SCR/.generated"
fi
OBJ=$(mktemp -d)
trap 'rm -rf "$OBJ"' EXIT
T=8
CF="-O3 -std=gnu99 -fcommon -w -fPIC -I$HERE/mpi_stub -I$SCR/src -DACTION_TYPE=TLSM -DNREPLICAS=1 \
 -DLOC_N0=$N0 -DLOC_N1=$N1 -DLOC_N2=$N2 -DLOC_N3=$N3 -DNRANKS_D3=$NR -DCOMMIT_HASH=oracle \
 -DDEODOETILE0=$T -DDEODOETILE1=$T -DDEODOETILE2=$T -DDEODOEGANG3=$T -DIMPSTAPTILE0=$T -DIMPSTAPTILE1=$T \
 -DIMPSTAPTILE2=$T -DIMPSTAPGANG3=$T -DSTAPTILE0=$T -DSTAPTILE1=$T -DSTAPTILE2=$T -DSTAPGANG3=$T \
 -DSIGMATILE0=$T -DSIGMATILE1=$T -DSIGMATILE2=$T -DSIGMAGANG3=$T"
SRCS="OpenAcc/fermion_matrix OpenAcc/sp_fermion_matrix OpenAcc/fermionic_utilities OpenAcc/sp_fermionic_utilities
 OpenAcc/inverter_multishift_full OpenAcc/sp_inverter_multishift_full OpenAcc/inverter_full OpenAcc/sp_inverter_full
 OpenAcc/inverter_mixedp OpenAcc/inverter_package OpenAcc/inverter_wrappers OpenAcc/float_double_conv OpenAcc/geometry
 OpenAcc/find_min_max OpenAcc/io OpenAcc/stouting OpenAcc/sp_stouting OpenAcc/plaquettes OpenAcc/sp_plaquettes OpenAcc/su3_utilities OpenAcc/sp_su3_utilities OpenAcc/fermion_force_utilities OpenAcc/sp_fermion_force_utilities OpenAcc/backfield OpenAcc/sp_backfield OpenAcc/backfield_parameters OpenAcc/fermion_force OpenAcc/sp_fermion_force OpenAcc/field_times_fermion_matrix Meas/ferm_meas
 RationalApprox/rationalapprox tests_and_benchmarks/test_and_benchmarks Mpi/multidev Mpi/communications Mpi/sp_communications Include/inverter_tricks"
pids=()
for f in $SRCS; do
  gcc $CF -c "$SCR/src/$f.c" -o "$OBJ/$(basename $f).o" & pids+=($!)
done
gcc $CF -c "$HERE/ref_shim.c" -o "$OBJ/ref_shim.o" & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
# -Bsymbolic: calls between reference functions bind inside this .so even if the product library (which
# exports the same names by design) is loaded in the same process
gcc -shared -Wl,-Bsymbolic -o "$OUT" "$OBJ"/*.o -lm
echo "built $OUT"
