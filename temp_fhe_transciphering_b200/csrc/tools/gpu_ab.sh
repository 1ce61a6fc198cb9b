set -x
timeout -s KILL 400 python temp_fhe_transciphering_b200/csrc/tools/max_dbg.py
