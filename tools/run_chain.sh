#!/bin/bash
# The shell recipe of the reference's test (test/ecoli_K12/run_test.sh:16-28: mkseqdb -> index x T_IDX -> overlap x T_OVLP ->
# cat | dedup) with the tools found in $BIN (default: this repository's GPU drop-ins; point it at oracle/_ref to run the
# unmodified reference).  `xargs -P` stands in for GNU parallel.
#   usage: tools/run_chain.sh <seq_dataset.lst> <workdir> [T_IDX=12] [T_OVLP=8] [JOBS=4]
set -e
LST=$1; WD=$2; TI=${3:-12}; TO=${4:-8}; JOBS=${5:-4}
BIN=${BIN:-$(cd "$(dirname "$0")/.." && pwd)/bin}
mkdir -p "$WD/index" "$WD/ovlp" "$WD/asm"
"$BIN/shmr_mkseqdb" -p "$WD/index/seq_dataset" -d "$LST" > "$WD/build_db.log" 2>&1
seq 1 "$TI" | xargs -P "$JOBS" -I{} sh -c "\"$BIN/shmr_index\" -p \"$WD/index/seq_dataset\" -r 6 -t $TI -c {} -o \"$WD/index/shmr\" > \"$WD/build_index.{}.log\" 2>&1"
seq -f "%02g" 1 "$TO" | xargs -P "$JOBS" -I{} sh -c "\"$BIN/shmr_overlap\" -p \"$WD/index/seq_dataset\" -l \"$WD/index/shmr-L2\" -t $TO -c {} -o \"$WD/ovlp/ovlp.{}\" 2> \"$WD/ovlp.{}.log\""
cat "$WD"/ovlp/ovlp.* | "$BIN/shmr_dedup" > "$WD/asm/preads.ovl"
echo "-" >> "$WD/asm/preads.ovl"
wc -l "$WD/asm/preads.ovl"
