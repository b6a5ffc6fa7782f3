// oracle/ref_shim — stand-in for <NazaraUtils/Bitset.hpp>
#pragma once
#include <NazaraUtils/Prerequisites.hpp>
#include <vector>

namespace Nz
{
	template<typename Block = UInt32> class Bitset
	{
		public:
			class Bit
			{
				public:
					Bit(Block& block, Block mask) : m_block(block), m_mask(mask) {}
					operator bool() const { return (m_block & m_mask) != 0; }
					Bit& operator=(bool v)
					{
						if (v)
							m_block |= m_mask;
						else
							m_block &= ~m_mask;
						return *this;
					}

				private:
					Block& m_block;
					Block m_mask;
			};

			void Clear() { m_blocks.clear(); m_bitCount = 0; }
			void Resize(std::size_t bitCount, bool defaultVal = false)
			{
				const std::size_t old = m_bitCount;
				m_blocks.resize((bitCount + BitsPerBlock - 1) / BitsPerBlock, defaultVal ? ~Block(0) : Block(0));
				m_bitCount = bitCount;
				if (defaultVal)
					for (std::size_t i = old; i < bitCount; ++i)
						(*this)[i] = true;
			}
			std::size_t GetSize() const { return m_bitCount; }
			bool Test(std::size_t bit) const { return (m_blocks[bit / BitsPerBlock] >> (bit % BitsPerBlock)) & Block(1); }
			bool operator[](std::size_t bit) const { return Test(bit); }
			Bit operator[](std::size_t bit) { return Bit(m_blocks[bit / BitsPerBlock], Block(1) << (bit % BitsPerBlock)); }

		private:
			static constexpr std::size_t BitsPerBlock = sizeof(Block) * 8;
			std::vector<Block> m_blocks;
			std::size_t m_bitCount = 0;
	};
}
