// Stand-in for the header CMakeRC generates at configure time (dependencies/tiny-cuda-nn/dependencies/cmrc/CMakeRC.cmake): the
// reference's testbed_nerf.cu only names the embedded file system when it JIT-compiles kernels, which oracle/ref never does.
// Test infrastructure (oracle/ref/ref_nerf_harness.cu); not reference code.
#pragma once
namespace cmrc { class embedded_filesystem {}; }
#define CMRC_DECLARE(ns) namespace cmrc { namespace ns { inline cmrc::embedded_filesystem get_filesystem() { return {}; } } }
