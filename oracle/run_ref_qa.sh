#!/bin/bash
# oracle/run_ref_qa.sh -- TEST INFRASTRUCTURE ONLY.
# Runs the reference's own CMC QA harness (tests/bin/run_cmc_check -> runQaTests.pl, Perl) for the
# BSIM4 and BSIM3 device tests against oracle/_ref/ngspice, in a scratch copy of the test
# directories (the reference tree is read-only), and prints the MATCH / DIFFER tally per suite.
# It pins the hand-built oracle binary (hand-written config.h and parsers) against the golden
# `reference/*.standard` vectors the reference ships for these models.
# Usage: oracle/run_ref_qa.sh [REF=/root/reference] [scratch=/tmp/ngb_qa]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"; W="${2:-/tmp/ngb_qa}"
[ -x "$HERE/_ref/ngspice" ] || { echo "oracle/_ref/ngspice not built"; exit 1; }
rm -rf "$W"; mkdir -p "$W/tests" "$W/sim"
cp -r "$REF/tests/bin" "$W/tests/bin"; cp -r "$REF/tests/bsim4" "$W/tests/bsim4"; cp -r "$REF/tests/bsim3" "$W/tests/bsim3"
chmod -R u+w "$W"; ln -s "$HERE/_ref/ngspice" "$W/sim/ngspice"
for suite in bsim4/nmos bsim4/pmos bsim3/nmos bsim3/pmos; do
  ( cd "$W/tests/$suite" && PATH="$W/sim:$PATH" sh ../../bin/run_cmc_check ngspice > "$W/$(echo $suite | tr / _).log" 2>&1 ) || true
  L="$W/$(echo $suite | tr / _).log"
  echo "$suite: standard-vs-reference MATCH $(grep -c 'compared to: reference) MATCH' "$L") DIFFER $(grep -c 'compared to: reference) DIFFER' "$L")" \
       "worst $(grep 'compared to: reference) DIFFER' "$L" | sed 's/.*error is //; s/%)//' | sort -g | tail -1)% ;" \
       "variants-vs-standard MATCH $(grep -c 'compared to: standard ) MATCH' "$L") DIFFER $(grep -c 'compared to: standard ) DIFFER' "$L")"
done
