// oracle/ref_shim — stand-in for <Nazara/Core/Error.hpp>: included by Ship.inl, nothing of it is used by the chunk pipeline.
#pragma once
#include <cassert>
#ifndef NazaraAssert
#define NazaraAssert(cond, ...) assert(cond)
#endif
