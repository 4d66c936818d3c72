#!/bin/bash
# Per-object count of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st
# -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, mbarrier -> SYNCS, legacy mma.sync -> HMMA).  Usage: bash tools/sass_table.sh > profiles/rNN_sass_table.txt
echo "# cuobjdump -sass leopard_b200/lib/obj/*.o | grep -c <mnemonic>   ($(nvcc --version | tail -2 | head -1))"
for o in leopard_b200/lib/obj/*.o; do n=$(basename $o .o); s=$(cuobjdump -sass $o 2>/dev/null)
  printf "%-12s UTCHMMA %4d (.2CTA %3d)  UTMALDG %3d  UTMASTG %3d  LDTM %3d  STTM %3d  UTCBAR %3d  SYNCS %4d  HMMA %d  LDGSTS %3d\n" $n \
    $(echo "$s" | grep -c "UTCHMMA") $(echo "$s" | grep -c "UTCHMMA.2CTA") $(echo "$s" | grep -c "UTMALDG") $(echo "$s" | grep -c "UTMASTG") \
    $(echo "$s" | grep -c "LDTM") $(echo "$s" | grep -c "STTM") $(echo "$s" | grep -c "UTCBAR") $(echo "$s" | grep -c "SYNCS") \
    $(echo "$s" | grep -c " HMMA") $(echo "$s" | grep -c "LDGSTS")
done
s=$(cuobjdump -sass leopard_b200/lib/libleopard_b200.so 2>/dev/null)
echo "libleopard_b200.so total: UTCHMMA $(echo "$s" | grep -c UTCHMMA), UTMALDG $(echo "$s" | grep -c UTMALDG), UTMASTG $(echo "$s" | grep -c UTMASTG), LDTM $(echo "$s" | grep -c LDTM), HMMA $(echo "$s" | grep -c ' HMMA')"
