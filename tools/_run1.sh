L=sentencepiece_b200/lib
for v in A B C A B C; do cp $L/variant_$v.so $L/libspm_b200.so; echo "variant $v"; timeout 200 python tools/lane_ab.py bpe32k:en --variants "CR=1;CR=0" --reps 6 2>&1 | tail -2; done
