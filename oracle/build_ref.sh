#!/bin/bash
# Builds the UNMODIFIED reference engine (libs/simulator + libs/superlu +
# libs/calculon + libs/data + cephes) from the sources where they lie under
# $REF (default /root/reference) into oracle/_ref/libeispice_ref.so.
# Test infrastructure only: used to pin the oracle restatement and as the CPU
# baseline ("kind": "reference").  Nothing is copied into the repo; objects and
# the .so go to oracle/_ref/ (git-ignored, but shipped to the GPU box).
# Recipe follows SURVEY.md section 8c.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
if [ ! -d "$REF/libs/simulator" ]; then
	echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)"
	exit 0
fi
mkdir -p "$OBJ"
INC="-I$REF/libs/log -I$REF/libs/data -I$REF/libs/calculon -I$REF/libs/superlu -I$REF/libs/simulator/include"
CF="-O2 -fPIC -w"
objs=""
cc() { # std, name, src, extra
	local o="$OBJ/$2.o"
	if [ ! -f "$o" ] || [ "$3" -nt "$o" ]; then
		gcc $CF -std=$1 $INC $4 -c "$3" -o "$o"
	fi
	objs="$objs $o"
}
for f in data dblhash hash list; do cc gnu99 data_$f "$REF/libs/data/$f.c"; done
for f in calc parser tokenizer; do cc gnu99 calc_$f "$REF/libs/calculon/$f.c"; done
SLU="dgssv dgssvx dsp_blas2 dsp_blas3 dgscon dlangs dgsequ dlaqgs dpivotgrowth dgsrfs dgstrf dgstrs dcopy_to_ucol dsnode_dfs dsnode_bmod dpanel_dfs dpanel_bmod dreadhb dcolumn_dfs dcolumn_bmod dpivotL dpruneL dmemory dutil dmyblas2 superlu_timer util memory get_perm_c mmd sp_coletree sp_preorder sp_ienv relax_snode heap_relax_snode colamd dlacon dlamch"
for f in $SLU; do cc gnu89 slu_$f "$REF/libs/superlu/SRC/$f.c" "-DAdd_"; done
for f in "$REF"/libs/netlib/cephes/*.c; do cc gnu99 ceph_$(basename $f .c) "$f"; done
for d in core devices math; do
	for f in "$REF"/libs/simulator/$d/*.c; do
		cc gnu99 sim_${d}_$(basename $f .c) "$f" "-DLIBNAME=simulator"
	done
done
cc gnu99 ref_shim "$HERE/ref_shim.c"
# link-time probe of SuperLU's permutations (ref_probe.c): the reference's own
# call of dgssvx is routed through it, no reference source is touched
cc gnu99 ref_probe "$HERE/ref_probe.c"
gcc -shared -Wl,--wrap=dgssvx -o "$OUT/libeispice_ref.so" $objs -lm
echo "build_ref: wrote $OUT/libeispice_ref.so"
