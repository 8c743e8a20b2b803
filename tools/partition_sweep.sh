#!/bin/bash
# SM-partition (green context) sweep for the chain-bound phase of dgetrf / dpotrf (device-resident, tools/lu_probe.py, tools/chol_probe.py)
echo "== baseline (no partition)"
LLACUDA_NO_PARTITION=1 python tools/lu_probe.py 2>&1 | grep dgetrf
LLACUDA_NO_PARTITION=1 python tools/chol_probe.py 2>&1 | grep dpotrf
for sms in 32 24 16; do
  for pm in 8192 10240; do
    echo "== LU  PANEL_SMS=$sms LU_PART_M=$pm"
    LLACUDA_PANEL_SMS=$sms LLACUDA_LU_PART_M=$pm python tools/lu_probe.py 2>&1 | grep "dgetrf"
  done
  for pn in 8192 10240 11264 12288 14336; do
    echo "== CHOL PANEL_SMS=$sms CHOL_PART_N=$pn"
    LLACUDA_PANEL_SMS=$sms LLACUDA_CHOL_PART_N=$pn python tools/chol_probe.py 2>&1 | grep "dpotrf"
  done
done
