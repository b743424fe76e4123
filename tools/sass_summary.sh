#!/bin/bash
# Opcode census of the shipped library (no GPU needed): the tcgen05 / TMA / TMEM instructions per kernel.
#   bash tools/sass_summary.sh > profiles/r02_sass_summary.txt
so=${1:-unet.cu_b200/lib/libunet_b200.so}
echo "# cuobjdump -sass $so  ($(git rev-parse --short HEAD 2>/dev/null), $(date -u +%F))"
cuobjdump -sass "$so" > /tmp/ub_sass.txt
echo "# whole library"
for op in UTCHMMA UTMALDG UTMASTG UTMAREDG UTMAPF UBLKPF LDTM STTM UTCBAR UTCATOMSWS "SYNCS" "ACQBULK" "UCGABAR" "MUFU.TANH" "MUFU.EX2" "RED.E" "REDG" "ATOMG"; do
  printf "%-12s %6d\n" "$op" "$(grep -c "$op" /tmp/ub_sass.txt)"
done
echo "# per kernel (kernels that issue tensor-core / TMA instructions)"
awk '/Function : /{name=$3} /UTCHMMA/{m[name]++} /UTMALDG/{l[name]++} /UTMASTG/{s[name]++} /UTMAREDG/{r[name]++} /LDTM/{t[name]++} /UTCBAR/{b[name]++} /UBLKPF/{p[name]++}
     END{for (n in m) printf "%-70s UTCHMMA %3d UTMALDG %3d UTMASTG %3d UTMAREDG %3d LDTM %3d UTCBAR %3d UBLKPF %3d\n", n, m[n], l[n], s[n], r[n], t[n], b[n], p[n]}' /tmp/ub_sass.txt | c++filt | sort
echo "# registers / shared memory per kernel (cuobjdump -res-usage)"
cuobjdump -res-usage "$so" 2>/dev/null | awk '/Function/{f=$2} /REG:/{print f, $1, $2, $3, $4}' | c++filt | grep -E "igemm|attn_tc|gn_|adamw|pack_weights|smallc|conv_out" | sort
