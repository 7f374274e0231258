#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the reference's OWN test programs (src/tests_and_benchmarks/deo_doe_test.c, inverter_multishift_test.c), unmodified,
# for one compile-time geometry, twice:
#   oracle/_ref/<prog>_ref_<geom>     every source of the reference (its gcc recipe: OpenACC pragmas ignored, CPU)
#   oracle/_ref/<prog>_staple_<geom>  the same main and host code, with the object files of the subsystems libstaple_b200.so
#                                     replaces LEFT OUT and the library linked instead; the only
#                                     source that is not the reference's is openstaple_b200/host/memory_wrapper_staple.c, which stands in for
#                                     Include/memory_wrapper.c (the allocation choke point)
# The second binary is the drop-in claim made executable: the reference's host program, its parser, generators and file
# writers, running its hot path on the B200 through the C ABI.  Both travel to the GPU box as binaries (oracle/_ref is
# git-ignored, not gpurun-ignored); tests/test_gpu_reference_host.py runs them there and compares their output files.
# With a fifth argument NRANKS_D3 > 1 the programs are built for that many D3 slabs against oracle/mpi_mini (a minimal MPI over
# local sockets: the image has none) and get the suffix _rN; run them with oracle/mpi_mini/mpirun.py -n N.
# usage: oracle/build_ref_host.sh N0 N1 N2 N3 [NRANKS_D3=1]
set -euo pipefail
REF=${STAPLE_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
N0=$1; N1=$2; N2=$3; N3=$4; NR=${5:-1}
GEOM=${N0}x${N1}x${N2}x${N3}
MPIDIR=$HERE/mpi_stub; MPISRC=$HERE/mpi_stub/mpi_single.c
if [ "$NR" -gt 1 ]; then GEOM=${GEOM}_r$NR; MPIDIR=$HERE/mpi_mini; MPISRC=$HERE/mpi_mini/mpi_mini.c; fi
mkdir -p "$HERE/_ref"
[ -d "$REF/src" ] || { echo "reference not present at $REF (prebuilt oracle/_ref is used as is)"; exit 0; }
"$HERE/build_ref.sh" $N0 $N1 $N2 $N3 $NR > /dev/null        # makes sure the scratch copy with the generated sp_* files exists
SCR=${STAPLE_ORACLE_SCRATCH:-${TMPDIR:-/tmp}/staple_oracle_src}
LIBDIR=$(cd "$HERE/../openstaple_b200" && pwd)
[ -f "$LIBDIR/libstaple_b200.so" ] || { echo "build libstaple_b200.so first"; exit 1; }
PROGLIST=${PROGS:-deo_doe_test inverter_multishift_test main}
STAMP="$HERE/_ref/${PROGLIST##* }_staple_$GEOM"          # the program linked last
if [ -f "$STAMP" ] && [ "$STAMP" -nt "$MPISRC" ] && [ "$STAMP" -nt "$HERE/../openstaple_b200/host/memory_wrapper_staple.c" ] && [ "$STAMP" -nt "$HERE/../openstaple_b200/host/multidev_staple.c" ] && [ "$STAMP" -nt "$0" ] && [ "$STAMP" -nt "$LIBDIR/../include/staple_b200.h" ]; then echo "up to date: $STAMP"; exit 0; fi
OBJ=$(mktemp -d)
trap 'rm -rf "$OBJ"' EXIT
T=8
CF="-O3 -std=gnu99 -fcommon -w -I$MPIDIR -I$SCR/src -DACTION_TYPE=TLSM -DNREPLICAS=1 \
 -DLOC_N0=$N0 -DLOC_N1=$N1 -DLOC_N2=$N2 -DLOC_N3=$N3 -DNRANKS_D3=$NR -DCOMMIT_HASH=oracle \
 -DDEODOETILE0=$T -DDEODOETILE1=$T -DDEODOETILE2=$T -DDEODOEGANG3=$T -DIMPSTAPTILE0=$T -DIMPSTAPTILE1=$T \
 -DIMPSTAPTILE2=$T -DIMPSTAPGANG3=$T -DSTAPTILE0=$T -DSTAPTILE1=$T -DSTAPTILE2=$T -DSTAPGANG3=$T \
 -DSIGMATILE0=$T -DSIGMATILE1=$T -DSIGMATILE2=$T -DSIGMAGANG3=$T"
# src/Makefile.am:47-172 (__common_sources) and :3-45 (__common_generated_sources)
read -r -d '' COMMON <<'LIST' || true
DbgTools/dbgtools Include/acceptances_info Include/debug Include/fermion_parameters Include/hash Include/inverter_tricks
Include/montecarlo_parameters Include/rep_info Include/setting_file_parser Include/tell_geom_defines Include/memory_wrapper
Meas/baryon_number_utilities Meas/ferm_meas Meas/gauge_meas Meas/magnetic_susceptibility_utilities Meas/polyakov Meas/measure_topo
Mpi/communications Mpi/multidev OpenAcc/HPT_utilities OpenAcc/action OpenAcc/alloc_settings OpenAcc/alloc_vars OpenAcc/backfield
OpenAcc/backfield_parameters OpenAcc/cooling OpenAcc/deviceinit OpenAcc/fermion_force OpenAcc/fermion_force_utilities
OpenAcc/fermionic_utilities OpenAcc/fermion_matrix OpenAcc/field_times_fermion_matrix OpenAcc/find_min_max OpenAcc/float_double_conv
OpenAcc/geometry OpenAcc/inverter_full OpenAcc/inverter_mixedp OpenAcc/inverter_multishift_full OpenAcc/inverter_package
OpenAcc/inverter_wrappers OpenAcc/io OpenAcc/ipdot_gauge OpenAcc/md_integrator OpenAcc/md_parameters OpenAcc/plaquettes
OpenAcc/random_assignement OpenAcc/rectangles OpenAcc/sp_fermion_force OpenAcc/stouting OpenAcc/su3_measurements OpenAcc/su3_utilities
OpenAcc/topological_action OpenAcc/topological_force OpenAcc/update_versatile Rand/random RationalApprox/rationalapprox
tests_and_benchmarks/test_and_benchmarks
OpenAcc/sp_alloc_vars OpenAcc/sp_backfield OpenAcc/sp_fermion_force_utilities OpenAcc/sp_fermionic_utilities OpenAcc/sp_fermion_matrix
OpenAcc/sp_inverter_full OpenAcc/sp_inverter_multishift_full OpenAcc/sp_ipdot_gauge OpenAcc/sp_md_integrator OpenAcc/sp_plaquettes
OpenAcc/sp_rectangles OpenAcc/sp_stouting OpenAcc/sp_su3_measurements OpenAcc/sp_su3_utilities OpenAcc/sp_topological_action
OpenAcc/sp_topological_force DbgTools/sp_dbgtools Meas/sp_gauge_meas Mpi/sp_communications
LIST
# the subsystems the library replaces whole.  plaquettes / su3_utilities / ferm_meas stay: the few
# functions of theirs that the library also exports are then defined twice, and the executable's own copy is the one its
# host code calls -- exactly what happens in a maintainer's build.
REPLACED=" Include/memory_wrapper OpenAcc/fermion_matrix OpenAcc/sp_fermion_matrix OpenAcc/fermionic_utilities OpenAcc/sp_fermionic_utilities
 OpenAcc/inverter_multishift_full OpenAcc/sp_inverter_multishift_full OpenAcc/inverter_full OpenAcc/sp_inverter_full OpenAcc/inverter_mixedp
 OpenAcc/inverter_package OpenAcc/inverter_wrappers OpenAcc/float_double_conv OpenAcc/find_min_max OpenAcc/fermion_force OpenAcc/sp_fermion_force
 OpenAcc/fermion_force_utilities OpenAcc/sp_fermion_force_utilities OpenAcc/field_times_fermion_matrix OpenAcc/stouting OpenAcc/sp_stouting "
pids=()
# DbgTools/debugger_hook.c is main's libdbghook.a (src/Makefile.am:176,187,216)
for f in $COMMON tests_and_benchmarks/deo_doe_test tests_and_benchmarks/inverter_multishift_test OpenAcc/main DbgTools/debugger_hook; do
  gcc $CF -c "$SCR/src/$f.c" -o "$OBJ/$(echo $f | tr / _).o" & pids+=($!)
done
gcc -O2 -std=gnu99 -w -I"$MPIDIR" -c "$MPISRC" -o "$OBJ/mpi_single.o" & pids+=($!)
gcc -O2 -std=gnu99 -w -I"$MPIDIR" -I"$HERE/../include" -DNRANKS_D3=$NR -DLOC_N0=$N0 -DLOC_N1=$N1 -DLOC_N2=$N2 -DLOC_N3=$N3 -c "$HERE/../openstaple_b200/host/memory_wrapper_staple.c" -o "$OBJ/host_shim.o" & pids+=($!)
# NR > 1 only: a third build, <prog>_staplemd_<geom>, also leaves src/Mpi/multidev.c out and takes openstaple_b200/host/multidev_staple.c
# (devinfo, pre_init_multidev1D, init_multidev1D, shutdown_multidev compiled against the host's mpi.h) in its place
if [ "$NR" -gt 1 ]; then gcc $CF -c "$HERE/../openstaple_b200/host/multidev_staple.c" -o "$OBJ/multidev_staple.o" & pids+=($!); fi
for p in "${pids[@]}"; do wait $p; done
ALL=""; KEPT=""
REPLACED=" $(echo $REPLACED) "     # one space between names, whatever the line breaks above
for f in $COMMON; do
  o="$OBJ/$(echo $f | tr / _).o"; ALL="$ALL $o"
  case "$REPLACED" in *" $f "*) ;; *) KEPT="$KEPT $o";; esac
done
# the third program is the reference's production main (OpenAcc/main.c: the whole RHMC), same two ways
for prog in $PROGLIST; do      # PROGS="deo_doe_test" builds a subset
  mo="$OBJ/tests_and_benchmarks_$prog.o"; [ $prog = main ] && mo="$OBJ/OpenAcc_main.o $OBJ/DbgTools_debugger_hook.o"
  gcc -o "$HERE/_ref/${prog}_ref_$GEOM" $mo $ALL "$OBJ/mpi_single.o" -lm
  gcc -o "$HERE/_ref/${prog}_staple_$GEOM" $mo $KEPT "$OBJ/host_shim.o" "$OBJ/mpi_single.o" \
      -L"$LIBDIR" -lstaple_b200 -Wl,-rpath,'$ORIGIN/../../openstaple_b200' -lm
  if [ "$NR" -gt 1 ]; then
    gcc -o "$HERE/_ref/${prog}_staplemd_$GEOM" $mo ${KEPT/$OBJ\/Mpi_multidev.o/} "$OBJ/multidev_staple.o" "$OBJ/host_shim.o" "$OBJ/mpi_single.o" \
        -L"$LIBDIR" -lstaple_b200 -Wl,-rpath,'$ORIGIN/../../openstaple_b200' -lm
  fi
done
echo "built $HERE/_ref/{${PROGS:-deo_doe_test,inverter_multishift_test,main}}_{ref,staple}_$GEOM"
