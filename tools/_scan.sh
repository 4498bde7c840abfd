for v in antmb7 antmb9; do GUARDX_LIB=/root/repo/guardx_b200/variants/libguardx_$v.so python tools/bench_art_quick.py ant 65536 40; done
python tools/bench_art_quick.py ant 65536 40
