python scripts/time_lhs.py 4096 2>&1 | grep TIMES
export VK_EMIT_JAC_TB=256 VK_EMIT_JAC_BLOCKS=1 VK_TAG=nosync
python -m vulcan_b200.build > /dev/null 2>&1; python scripts/time_lhs.py 4096 2>&1 | grep TIMES
export VK_TAG=rowsync
sed -i 's/^#define VK_EMITJ_ROW_SYNC$/#define VK_EMITJ_ROW_SYNC __syncthreads();/' vulcan_b200/csrc/vk_emit_rt.cuh
python -m vulcan_b200.build > /dev/null 2>&1; python scripts/time_lhs.py 4096 2>&1 | grep TIMES
export VK_EMIT_JAC_TB=128 VK_EMIT_JAC_BLOCKS=2 VK_TAG=rowsync128
python -m vulcan_b200.build > /dev/null 2>&1; python scripts/time_lhs.py 4096 2>&1 | grep TIMES
