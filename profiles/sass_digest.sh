#!/bin/bash
# SASS evidence per kernel of the built library: the Blackwell-specific mnemonics (tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM,
# cp.async.bulk -> UBLKCP, mbarrier -> SYNCS) and the arithmetic mix.  bash profiles/sass_digest.sh > profiles/r2c_sass_digest.txt
so=${1:-hmmufotu_b200/lib/libhmmufotu_b200.so}
cuobjdump -sass $so | awk '
/Function :/ { name=$3; sub(/^_ZN2hu[0-9]+/, "", name); fn[name]=1; cur=name; next }
/^[ \t]+\/\*[0-9a-f]+\*\// { op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); cnt[cur, op]++; tot[cur]++ }
END {
  n=split("UTCIMMA UTCHMMA UTCBAR LDTM STTM UBLKCP UTMALDG SYNCS DFMA DADD DMUL DSETP MUFU POPC IDP IMAD SHFL LDS STS LDG STG BAR", ops, " ");
  for (f in fn) { line=sprintf("%-70s total %6d |", substr(f,1,70), tot[f]); for (i=1;i<=n;i++) if (cnt[f, ops[i]]) line=line sprintf(" %s %d", ops[i], cnt[f, ops[i]]); print line }
}' | sort
