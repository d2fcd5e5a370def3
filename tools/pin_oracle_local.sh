#!/bin/bash
# Pins the CPU oracle's usearch_local path against the UNMODIFIED reference binary.
#   tools/pin_oracle_local.sh Q.fa DB.fa ID EVALUE aa|nt [maxaccepts maxrejects]
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
Q=$1; DB=$2; ID=$3; EV=$4; AL=$5; MA=${6:-}; MR=${7:-}
T=$(mktemp -d)
EXTRA=""
[ -n "$MA" ] && EXTRA="-maxaccepts $MA -maxrejects $MR"
[ "$AL" = nt ] && EXTRA="$EXTRA -strand plus"
$HERE/oracle/_ref/usearch12 -usearch_local $Q -db $DB -id $ID -evalue $EV -threads 8 $EXTRA \
  -uc $T/r.uc -blast6out $T/r.b6 -userout $T/r.user \
  -userfields query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+evalue+bits+raw+caln+qstrand -quiet
$HERE/oracle/_build/uso_cli usearch_local $Q $DB $ID $EV $AL $T/o.user $T/o.uc $T/o.b6 $MA $MR
rc=0
for x in user uc b6; do
  if cmp -s <(sort $T/r.$x) <(sort $T/o.$x); then echo "IDENTICAL $x ($(wc -l < $T/r.$x) lines)"; else echo "DIFF $x"; diff <(sort $T/r.$x) <(sort $T/o.$x) | head -6; rc=1; fi
done
rm -rf $T
exit $rc
