#!/bin/bash
# First GPU call of the next round: everything round 1 wrote after its GPU budget was spent.
# Usage (from the repo root, on the GPU box):  bash tools/next_round_gpu.sh   -> logs under gpurun_out/
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
# 1. the file -> PCM chain through the verified entry points (front-ends on the CPU), incl. independent per-channel block types
timeout 600 python -m pytest tests/test_zz_file_to_pcm.py -m gpu -q > gpurun_out/chain.log 2>&1; tail -3 gpurun_out/chain.log
# 2. FLAC restoration kernel after the FIXED-order-1 fix, and the .flac file chain
SYMGPU_TEST_FLAC=1 timeout 600 python -m pytest tests/test_flac_parity_gpu.py tests/test_flac_frontend.py -m gpu -q > gpurun_out/flac.log 2>&1; tail -3 gpurun_out/flac.log
# 3. the device entropy front-end (never run before): under the sanitizer first, then plain
SYMGPU_TEST_ENTROPY=1 timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_mp3_entropy_gpu.py -x -q > gpurun_out/entropy_memcheck.log 2>&1; tail -5 gpurun_out/entropy_memcheck.log
SYMGPU_TEST_ENTROPY=1 timeout 600 python -m pytest tests/test_mp3_entropy_gpu.py -q > gpurun_out/entropy.log 2>&1; tail -3 gpurun_out/entropy.log
# 4. end to end from file bytes, both front-ends
timeout 600 python tools/mp3_file_e2e_bench.py > gpurun_out/mp3_file_e2e.json 2> gpurun_out/mp3_file_e2e.err; cat gpurun_out/mp3_file_e2e.json
# 5. Vorbis-in-Ogg file bytes -> PCM (front-end on the CPU, verified synthesis + output stage on the GPU)
SYMGPU_TEST_VORBIS_CHAIN=1 timeout 600 python -m pytest tests/test_zz_ogg_vorbis_to_pcm.py -m gpu -q > gpurun_out/vorbis_chain.log 2>&1; tail -3 gpurun_out/vorbis_chain.log
# 6. ADTS file bytes -> PCM (AAC-LC front-end on the CPU, verified synthesis + output stage on the GPU)
SYMGPU_TEST_AAC_CHAIN=1 timeout 600 python -m pytest tests/test_zz_adts_aac_to_pcm.py -m gpu -q > gpurun_out/aac_chain.log 2>&1; tail -3 gpurun_out/aac_chain.log
# 7. many files of all three codecs at once: one synthesis launch per codec
SYMGPU_TEST_MANY_FILES=1 timeout 600 python -m pytest tests/test_zz_many_files.py -m gpu -q > gpurun_out/many_files.log 2>&1; tail -3 gpurun_out/many_files.log
timeout 900 python tools/files_e2e_bench.py > gpurun_out/files_e2e.json 2> gpurun_out/files_e2e.err; cat gpurun_out/files_e2e.json
