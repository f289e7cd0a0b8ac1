#!/bin/bash
# usage: tools/ncu_export.sh <report-base>   -> writes <base>.raw.csv and <base>.source.csv next to the .ncu-rep
set -e
base="$1"
ncu -i "$base.ncu-rep" --page raw --csv > "$base.raw.csv" 2>/dev/null || true
ncu -i "$base.ncu-rep" --page source --csv > "$base.source.csv" 2>/dev/null || true
