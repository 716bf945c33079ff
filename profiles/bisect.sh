for v in v1 v2 v3; do echo "== $v"; LS_LIB=$PWD/minimax-speech_b200/libls_$v.so timeout 120 python profiles/debug_fail.py 2>&1 | grep -E "ok|Error" | tail -3; done
echo "== default halo0"; LS_CONV_HALO=0 timeout 120 python profiles/debug_fail.py 2>&1 | grep -E "ok|Error" | tail -3
