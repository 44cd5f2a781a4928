for v in "" _p8 _p2; do echo "variant $v"; GAPLAC_B200_LIB=$PWD/gaplac_b200/libgaplac_b200$v.so timeout 120 python tools/run_c5.py 8192 3 2>&1 | tail -1; done
