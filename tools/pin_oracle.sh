#!/bin/bash
# Pins the CPU oracle (oracle/uso.c) against the UNMODIFIED reference binary (oracle/_ref/usearch12)
# at full config sizes: runs both on the same FASTA inputs and diffs the sorted output files.
#   tools/pin_oracle.sh Q.fa DB.fa ID plus|both [maxaccepts maxrejects]
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
Q=$1; DB=$2; ID=$3; STRAND=$4; MA=${5:-}; MR=${6:-}
T=$(mktemp -d)
EXTRA=""
[ -n "$MA" ] && EXTRA="-maxaccepts $MA -maxrejects $MR"
$HERE/oracle/_ref/usearch12 -usearch_global $Q -db $DB -id $ID -strand $STRAND -threads 8 $EXTRA \
  -uc $T/r.uc -blast6out $T/r.b6 -userout $T/r.user \
  -userfields query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand -quiet
$HERE/oracle/_build/uso_cli usearch_global $Q $DB $ID $STRAND $T/o.user $T/o.uc $T/o.b6 $MA $MR
rc=0
for x in user uc b6; do
  if cmp -s <(sort $T/r.$x) <(sort $T/o.$x); then echo "IDENTICAL $x ($(wc -l < $T/r.$x) lines)"; else echo "DIFF $x"; diff <(sort $T/r.$x) <(sort $T/o.$x) | head -6; rc=1; fi
done
rm -rf $T
exit $rc
