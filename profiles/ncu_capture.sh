#!/bin/bash
# usage: profiles/ncu_capture.sh TAG KERNEL_REGEX [skip] [driver args...]   (run on the GPU box)
# Captures ONE launch with --set full and keeps only the CSV pages (the .ncu-rep embeds the whole cubin: ~50 MB).
TAG=$1; RE=$2; SKIP=${3:-1}; shift 3
OUT=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RE" -s $SKIP -c 1 -o /tmp/prof_$TAG \
    python profiles/prof_driver.py --reps ${REPS:-4} "$@" > $OUT/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/ncu_${TAG}_sass.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page details > $OUT/ncu_${TAG}_details.txt 2>/dev/null
tail -1 $OUT/ncu_$TAG.log
