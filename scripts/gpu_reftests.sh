#!/bin/bash
# The SDK's own gtest suite (CPU part) linked against libomm-b200.so (oracle/Makefile target `reftests`), run from the repo root.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 ./oracle/_ref/tests_b200 --gtest_filter=-GpuTest.* --gtest_brief=1 > gpurun_out/reftests_b200.txt 2>&1
tail -25 gpurun_out/reftests_b200.txt
