// oracle/ref_shim — stand-in for <fmt/format.h>: the chunk pipeline only prints two diagnostics in block picking
#pragma once
namespace fmt
{
	template<typename... Args> void print(const char*, Args&&...) {}
}
