# strict micro-benchmark (tools/strict_microbench.cu): variants x chunk counts
for b in tools/sb_*; do case $b in *.cu|*.sh) continue;; esac
echo -n "$b auto: "; $b 100000 200
for c in 3 6; do echo -n "$b chunks=$c: "; GKB_NL_CHUNKS=$c $b 100000 200; done
echo -n "$b auto 1000 epochs: "; $b 100000 1000
done
