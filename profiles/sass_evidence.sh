#!/bin/bash
# SASS evidence for the fused update kernel (run on the CPU box; cuobjdump needs no GPU):
#   profiles/sass_evidence.sh > profiles/r1_d_sass_update_kernel.txt
cd "$(dirname "$0")/.."
LIB=bevy_firework_b200/libfirework_b200.so
cuobjdump -sass $LIB > /tmp/fw_all.sass
echo "# $(cuobjdump -lelf $LIB | tr '\n' ' ')"
awk '/Function : .*update_kernelILb0ELi0/{f=1} f&&/Function : /&&!/update_kernelILb0ELi0/{f=0} f' /tmp/fw_all.sass > /tmp/fw_upd.sass
echo "# FIFO update kernel (update_kernel<false,0>): $(grep -c '^ *\/\*[0-9a-f]\{4\}\*\/' /tmp/fw_upd.sass) SASS instructions"
echo "# (the pack pointers come out of a descriptor table, so the 128/64-bit pack accesses are generic LD.E/ST.E, not LDG/STG;"
echo "#  explicit ld.global/st.global variants measured no faster: profiles/r1_tuning.md, cache-operator row)"
for m in "LD.E.128" "LD.E.64" "ST.E.128" "ST.E.64" "LDG.E.128" "LDG.E.64" "STG.E.128" "STG.E.64" "STG.E " "UBLKCP" "SYNCS" "CREDUX" "REDG.E" "VOTE" "BAR.SYNC" "FFMA" "FMUL" "FADD" "MUFU" "LDS" "LD.E" "ST.E"; do
  printf "%-12s %s\n" "$m" "$(grep -c "$m" /tmp/fw_upd.sass)"
done
echo "# global memory instructions:"
grep -E "LDG|STG|REDG|UBLKCP| LD\.E| ST\.E" /tmp/fw_upd.sass | sed -e 's/^ *\/\*[0-9a-f]*\*\/ *//' -e 's/ *\/\*.*//' | awk '{print $1, $2}' | sort | uniq -c | sort -rn
