#!/bin/bash
# SASS / ptxas evidence for the built library (no GPU needed): per kernel family the TMA / bulk-copy / mbarrier / named-barrier
# instruction counts (cuobjdump -sass) and the register / spill table (cuobjdump -res-usage).
# Usage: scripts/sass_summary.sh > profiles/r02_sass_summary.txt
so=seismic_cpml_b200/libcpml_b200.so
echo "# $so : $(cuobjdump -lelf $so | head -3 | tr '\n' ' ')"
echo "# default kernels per workload: cfg3 / cfg4 k_{stress,velocity}3d_ws<double,1,104,8>; cfg3f <float,1,104,8>; cfg5 k_vstress3d<32,8,2,0> + k_vvelocity3d_ws<64,8> (1024-wide slabs; cfg5d, 210 wide: <108,8>)"
echo
printf "%-100s %8s %8s %8s %8s %8s %8s\n" "kernel (cuobjdump -sass)" UTMALDG UBLKCP SYNCS BAR.ARV BAR.SYNC STG.EF
for fn in $(cuobjdump -sass $so | grep -oE "Function : \S+" | awk '{print $3}' | grep -E "3d_ws|3d_tmaILb1ELi104ELi8ELi1|k_vstress3dILi32ELi8ELi2ELi0|k_vvelocity3dILi32ELi8ELi2|2d_pairILi4ELi32ELi8ELi2" | grep -vE "Lb0E" | sort -u); do
  cuobjdump -sass -fun "$fn" $so 2>/dev/null > /tmp/_k.sass
  printf "%-100s %8d %8d %8d %8d %8d %8d\n" "$(echo $fn | c++filt | cut -c1-100)" \
    $(grep -c UTMALDG /tmp/_k.sass) $(grep -c UBLKCP /tmp/_k.sass) $(grep -c "SYNCS" /tmp/_k.sass) $(grep -c "BAR.ARV" /tmp/_k.sass) $(grep -c "BAR.SYNC" /tmp/_k.sass) $(grep -c "STG.E.EF" /tmp/_k.sass)
done
echo
echo "# registers / spills / shared memory (cuobjdump -res-usage; STACK > 0 = spill frame)"
cuobjdump -res-usage $so 2>/dev/null | grep -A1 -E "Function .*(3d_ws|3d_tmaILb1ELi104ELi8ELi1|k_vstress3dILi32ELi8ELi2ELi0|k_vvelocity3dILi32ELi8ELi2|2d_pairILi4ELi32ELi8ELi2|vstress2d|vvelocity2d)" | grep -v "^--" | paste - - | sed 's/ Function //' | while read fn rest; do echo "$(echo ${fn%:} | c++filt | cut -c1-90) | $rest"; done | grep -v "Lb0E\|<(bool)0" 
