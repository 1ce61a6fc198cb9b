set -x
mkdir -p gpurun_out
CBS_TRACE_VARIANT=3 python temp_fhe_transciphering_b200/csrc/tools/trace_dbg.py gpurun_out/t3 want
CBS_TRACE_VARIANT=4 python temp_fhe_transciphering_b200/csrc/tools/trace_dbg.py gpurun_out/t4
