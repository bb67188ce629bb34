#!/bin/bash
MARBLER_B200_LIB=$PWD/marbler_b200/libmarbler_b200_trace.so python scripts/tc2_trace.py 2>&1 | tail -12
