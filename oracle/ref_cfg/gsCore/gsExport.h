// Hand-written stand-in for the cmake-generated gsCore/gsExport.h
// (from src/gsCore/gsExport.h.in): static, header-only consumption.
#pragma once
#include <gsCore/gsConfig.h>
#define GISMO_EXPORT
#define GISMO_IMPORT
#define GISMO_DEFAULT_VIS
#define STRUCT_TEMPLATE_INST   template struct
#define CLASS_TEMPLATE_INST    template class
#define TEMPLATE_INST          template
#define EXTERN_STRUCT_TEMPLATE extern template struct
#define EXTERN_CLASS_TEMPLATE  extern template class
#define EXTERN_TEMPLATE        extern template
#ifndef __has_feature
#define __has_feature(x) 0
#endif
#define GISMO_FINAL final
#define GISMO_OVERRIDE override
