// oracle/ref_shim — stand-in for <NazaraUtils/Flags.hpp>
#pragma once
#include <NazaraUtils/Prerequisites.hpp>

namespace Nz
{
	template<typename E> class Flags
	{
		public:
			using BitField = UInt32;
			static constexpr BitField ValueMask = (BitField(1) << (static_cast<BitField>(E::Max) + 1)) - 1;

			constexpr Flags(BitField value = 0) : m_value(value) {}
			constexpr Flags(E e) : m_value(BitField(1) << static_cast<BitField>(e)) {}
			constexpr bool Test(const Flags& f) const { return (m_value & f.m_value) == f.m_value; }
			constexpr explicit operator bool() const { return m_value != 0; }
			constexpr operator BitField() const { return m_value; }
			constexpr Flags& operator|=(const Flags& f) { m_value |= f.m_value; return *this; }
			constexpr Flags operator|(const Flags& f) const { return Flags(m_value | f.m_value); }
			constexpr Flags operator&(const Flags& f) const { return Flags(m_value & f.m_value); }

		private:
			BitField m_value;
	};
}
