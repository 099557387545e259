// launch.h -- every kernel launch of libsbmf_cuda goes through SBMF_LAUNCH: <<<grid, block, smem, stream>>> on the GPU.  Under
// SBMF_SIMT_EMU (a test-only host build of the same sources against tools/emu_include: the memory- and race-check of the kernels,
// tools/sbmf_sanitize.sh) the kernel body is executed CTA by CTA on host threads instead.  `kernel` is passed in parentheses
// because template argument lists contain commas.
#pragma once
#ifdef SBMF_SIMT_EMU
#define SBMF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    simt::launch((grid), (block), [&] { auto kfn_ = kernel; kfn_(__VA_ARGS__); }, (size_t)(smem))
#else
#define SBMF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    [&] { auto kfn_ = kernel; kfn_<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); }()
#endif
