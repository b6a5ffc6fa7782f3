// oracle/ref_shim — stand-in for <NazaraUtils/CallOnExit.hpp>
#pragma once
#include <utility>

namespace Nz
{
	template<typename F> class CallOnExit
	{
		public:
			CallOnExit(F&& f) : m_f(std::move(f)) {}
			CallOnExit(const CallOnExit&) = delete;
			~CallOnExit() { m_f(); }

		private:
			F m_f;
	};
}
#define NAZARA_SHIM_CAT2(a, b) a##b
#define NAZARA_SHIM_CAT(a, b) NAZARA_SHIM_CAT2(a, b)
#define NAZARA_DEFER(expr) Nz::CallOnExit NAZARA_SHIM_CAT(nzDefer, __LINE__)([&]() { expr })
