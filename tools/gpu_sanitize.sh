#!/bin/bash
# compute-sanitizer (one tool per call: TOOL=memcheck|racecheck|synccheck) on small parity cases
mkdir -p gpurun_out
TOOL=${TOOL:-memcheck}
SEL='test_stft_works_kat or test_real_to_complex_kat or (test_perform_stft_parity and (20000-320-80-512 or 30000-884-221-1024 or 60000-1920-480-2048 or 2048-2048-512-2048 or 90000-4096-256-4096 or 150000-8192)) or (test_default_mel_db_parity and (8000 or 48000)) or (test_fixed_mel_db_parity and (4096-256 or 256-64)) or test_multitrack_golden_clips or test_stereo_is_channel_sum or (test_grey_to_rgb_parity and (shape3 or shape7 or shape5)) or test_wav_image_parity'
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 77 --log-file gpurun_out/sanitizer_$TOOL.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 1200 -k "$SEL" > gpurun_out/sanitizer_${TOOL}_pytest.log 2>&1
echo "sanitizer $TOOL exit $?"
tail -3 gpurun_out/sanitizer_${TOOL}_pytest.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error" gpurun_out/sanitizer_$TOOL.log | head -10
