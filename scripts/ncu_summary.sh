#!/bin/bash
# usage: scripts/ncu_summary.sh gpurun_out/<report>.ncu-rep profiles/<name>
# writes <name>_details.txt (ncu --page details), <name>_raw.csv (--page raw) and <name>_sass_mix.txt (opcode mix + stall reasons)
set -e
rep=$1; out=$2
ncu -i "$rep" --page details > "${out}_details.txt" 2>/dev/null
ncu -i "$rep" --page raw --csv > "${out}_raw.csv" 2>/dev/null
ncu -i "$rep" --page source --csv --print-source sass > /tmp/_sass.csv 2>/dev/null
python "$(dirname "$0")/sass_mix.py" /tmp/_sass.csv > "${out}_sass_mix.txt"
grep -E "Duration|Registers Per Thread|Achieved Occupancy|Executed Ipc Active|Issue Slots Busy" "${out}_details.txt" | head -8
