#!/bin/bash
# profiles/r02_sass_tcgen05.txt: SASS evidence of the Blackwell-native instructions (cuobjdump -sass of the in-tree objects; no GPU needed)
OBJ=crnn-ocr-lite_b200/build/gemm_tc.o
strip() { sed 's/ *\/\* 0x[0-9a-f]* \*\///' | cut -c1-150; }
echo "# SASS evidence of the Blackwell-native instructions in crnn-ocr-lite_b200/libcrnn_b200.so (round 2; cuobjdump -sass of the in-tree objects, sm_100a; tools/sass_listing.sh)"; echo
echo "Mnemonic counts per object -- UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA), UTMALDG = cp.async.bulk.tensor (tensor-map TMA),"
echo "SYNCS = mbarrier ops, UTCATOMSWS = tcgen05.alloc/dealloc, HMMA = mma.sync (recurrence), ELECT = elect.sync, UCGABAR = barrier.cluster, MAPA = cluster address mapping:"; echo
for o in crnn-ocr-lite_b200/build/*.o; do
  cuobjdump -sass $o > /tmp/o.sass 2>/dev/null; line="$(basename $o):"
  for m in UTCHMMA LDTM UTCBAR UBLKCP UTMALDG SYNCS UTCATOMSWS ' HMMA' ELECT UCGABAR ' MAPA' 'ST.ASYNC\|STAS'; do c=$(grep -c -- "$m" /tmp/o.sass); [ "$c" != "0" ] && line="$line ${m# }=$c"; done
  echo "  $line"
done
echo; echo "## xw_gemm_tc_v2_kernel<false>: the tcgen05 / TMA / elect instructions in program order"; echo
cuobjdump -sass $OBJ | awk '/Function.*xw_gemm_tc_v2_kernelILb0/{p=1} /Function.*xw_gemm_tc_v2_kernelILb1/{p=0} p' | grep -E "UTCHMMA|LDTM|UTCBAR|UBLKCP|UTMALDG|UTCATOMSWS|ELECT|FENCE.VIEW.ASYNC" | strip
echo; echo "## xw_gemm_tc_v2_kernel<true> (CTA pair, cta_group::2, behind CRNN_GEMM_PAIR=1): the 2-CTA forms"; echo
cuobjdump -sass $OBJ | awk '/Function.*xw_gemm_tc_v2_kernelILb1/{p=1} /Function.*prep_weight|Function.*xty_gemm/{p=0} p' | grep -E "UTCHMMA|UTCBAR|UTCATOMSWS|UCGABAR|MAPA" | strip | head -40
echo; echo "## xty_gemm_tc_kernel<256> (weight gradients)"; echo
cuobjdump -sass $OBJ | awk '/Function.*xty_gemm_tc_kernelILi256/{p=1} /Function.*prep_weight/{p=0} p' | grep -E "UTCHMMA|LDTM|UTCBAR|UTCATOMSWS|RED.E" | strip | head -40
