// Hand-written stand-in for the header the reference's cmake step would
// generate from src/gsCore/gsConfig.h.in (reference file, not copied: only the
// macro *values* a default configuration selects are stated here).
// Test infrastructure only: lets oracle/Makefile compile the reference's own
// assembler sources, header-only, straight from /root/reference.
#pragma once
#define GISMO_VERSION "24.8.0"
#define GISMO_MAJOR 24
#define GISMO_MINOR 8
#define GISMO_PATCH 0
#define GISMO_COEFF_TYPE double
#ifndef real_t
#define real_t GISMO_COEFF_TYPE
#endif
#define index_t int
#define short_t int
#ifndef GISMO_DATA_DIR
#define GISMO_DATA_DIR "/root/reference/filedata/"
#endif
#define GISMO_SEARCH_PATHS GISMO_DATA_DIR
#define GISMO_CONFIG_DIR "/tmp/"
#include <gsCore/gsConfigExt.h>
// GISMO_BUILD_LIB deliberately NOT defined: pure template (header-only) build.
#define EIGEN_DEFAULT_DENSE_INDEX_TYPE index_t
#define EIGEN_DEFAULT_SPARSE_INDEX_TYPE index_t
#define EIGEN_DEFAULT_TO_COL_MAJOR
#define EIGEN_NO_STATIC_ASSERT
