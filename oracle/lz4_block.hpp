// TEST INFRASTRUCTURE — CPU restatement of the LZ4 block format as the reference uses it for `Packets::ChunkReset`
// (src/CommonLib/Protocol/Packets.cpp:172-212 -> src/CommonLib/Utility/BinaryCompressor.cpp:17-50):
//   Compress   = LZ4_compress_fast_extState(state, src, dst, size, LZ4_compressBound(size), acceleration 0 -> 1)
//   Decompress = LZ4_decompress_safe(src, dst, compressedSize, maxOutputSize)
// lz4 is a third-party dependency that is absent from /root/reference (xmake package "lz4", xmake.lua:25, no version pin).  This file
// restates the published algorithm of lz4 1.9.x / 1.10 (lib/lz4.c, LZ4_compress_generic with tableType byU16: the path every input
// shorter than 64 KiB + 11 bytes takes, which covers the 65,536-byte block array of a chunk) from the format description and the
// behaviour of the library; nothing here is copied.  PINNED: tests/test_lz4.py compares the compressor byte for byte with the LZ4
// library bundled in pyarrow (codec "lz4_raw" = LZ4_compress_default = the same function with acceleration 1) on chunk-shaped and random
// inputs, and decodes the library's streams with the decompressor below.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use anything under oracle/.
#pragma once
#include <cstdint>
#include <cstring>

namespace tsom_oracle_lz4
{
	constexpr int MINMATCH = 4;
	constexpr int MFLIMIT = 12;      // no match may start within the last 12 bytes
	constexpr int LASTLITERALS = 5;  // the last 5 bytes are always literals
	constexpr int MIN_LENGTH = MFLIMIT + 1;
	constexpr int HASHLOG_U16 = 13;  // LZ4_MEMORY_USAGE 14 -> 16 KiB table -> 8192 16-bit entries
	constexpr int SKIP_TRIGGER = 6;
	constexpr int LIMIT_64K = 65536 + (MFLIMIT - 1);
	constexpr int MAX_INPUT = 0x7E000000;

	inline int compress_bound(int n) { return (unsigned)n > (unsigned)MAX_INPUT ? 0 : n + n / 255 + 16; }

	inline std::uint32_t read32(const std::uint8_t* p)
	{
		std::uint32_t v;
		std::memcpy(&v, p, 4);
		return v;
	}

	inline std::uint32_t hash_u16(std::uint32_t sequence) { return (sequence * 2654435761u) >> (MINMATCH * 8 - HASHLOG_U16); }

	// number of equal bytes of in / match, never reading `in` at or past limit
	inline unsigned common_length(const std::uint8_t* in, const std::uint8_t* match, const std::uint8_t* limit)
	{
		const std::uint8_t* start = in;
		while (in < limit && *in == *match)
		{
			++in;
			++match;
		}
		return unsigned(in - start);
	}

	// -> compressed size, 0 on failure (input too long for this restatement, or dst too small: the reference always passes the bound)
	inline int compress_block(const std::uint8_t* src, int n, std::uint8_t* dst, int cap)
	{
		if (n < 0 || n >= LIMIT_64K || cap < compress_bound(n))
			return 0;
		std::uint16_t table[1 << HASHLOG_U16];
		std::memset(table, 0, sizeof(table)); // a fresh state: every entry points at offset 0

		const std::uint8_t* ip = src;
		const std::uint8_t* anchor = src;
		const std::uint8_t* const iend = src + n;
		const std::uint8_t* const mflimitPlusOne = iend - MFLIMIT + 1;
		const std::uint8_t* const matchlimit = iend - LASTLITERALS;
		std::uint8_t* op = dst;

		if (n >= MIN_LENGTH)
		{
			table[hash_u16(read32(ip))] = 0; // the first position
			++ip;
			std::uint32_t forwardH = hash_u16(read32(ip));
			for (;;)
			{
				const std::uint8_t* match;
				std::uint8_t* token;
				// find a match: probe the table at a step that grows by one every 64 misses
				{
					const std::uint8_t* forwardIp = ip;
					int step = 1;
					int searchMatchNb = 1 << SKIP_TRIGGER;
					bool found = false;
					for (;;)
					{
						const std::uint32_t h = forwardH;
						const std::uint32_t current = std::uint32_t(forwardIp - src);
						const std::uint32_t matchIndex = table[h];
						ip = forwardIp;
						forwardIp += step;
						step = searchMatchNb++ >> SKIP_TRIGGER;
						if (forwardIp > mflimitPlusOne)
							break; // -> last literals
						match = src + matchIndex;
						forwardH = hash_u16(read32(forwardIp));
						table[h] = std::uint16_t(current);
						if (read32(match) == read32(ip))
						{
							found = true;
							break;
						}
					}
					if (!found)
						break;
				}
				// catch up: extend the match backwards over the pending literals
				while (ip > anchor && match > src && ip[-1] == match[-1])
				{
					--ip;
					--match;
				}
				// literals
				{
					const unsigned litLength = unsigned(ip - anchor);
					token = op++;
					if (litLength >= 15)
					{
						int len = int(litLength) - 15;
						*token = 15 << 4;
						for (; len >= 255; len -= 255)
							*op++ = 255;
						*op++ = std::uint8_t(len);
					}
					else
						*token = std::uint8_t(litLength << 4);
					std::memcpy(op, anchor, litLength);
					op += litLength;
				}
				bool lastLiterals = false;
				for (;;) // one sequence per round; a second round when the position right after a match matches again
				{
					// offset
					const unsigned offset = unsigned(ip - match);
					*op++ = std::uint8_t(offset);
					*op++ = std::uint8_t(offset >> 8);
					// match length
					{
						unsigned matchCode = common_length(ip + MINMATCH, match + MINMATCH, matchlimit);
						ip += matchCode + MINMATCH;
						if (matchCode >= 15)
						{
							*token += 15;
							matchCode -= 15;
							for (; matchCode >= 255; matchCode -= 255)
								*op++ = 255;
							*op++ = std::uint8_t(matchCode);
						}
						else
							*token += std::uint8_t(matchCode);
					}
					anchor = ip;
					if (ip >= mflimitPlusOne)
					{
						lastLiterals = true;
						break;
					}
					// fill the table at ip - 2, then test the new position at once
					table[hash_u16(read32(ip - 2))] = std::uint16_t(ip - 2 - src);
					const std::uint32_t h = hash_u16(read32(ip));
					const std::uint32_t current = std::uint32_t(ip - src);
					match = src + table[h];
					table[h] = std::uint16_t(current);
					if (read32(match) == read32(ip))
					{
						token = op++;
						*token = 0;
						continue;
					}
					break;
				}
				if (lastLiterals)
					break;
				forwardH = hash_u16(read32(++ip));
			}
		}
		// last literals
		{
			const std::size_t lastRun = std::size_t(iend - anchor);
			if (lastRun >= 15)
			{
				std::size_t acc = lastRun - 15;
				*op++ = 15 << 4;
				for (; acc >= 255; acc -= 255)
					*op++ = 255;
				*op++ = std::uint8_t(acc);
			}
			else
				*op++ = std::uint8_t(lastRun << 4);
			std::memcpy(op, anchor, lastRun);
			op += lastRun;
		}
		return int(op - dst);
	}

	// LZ4_decompress_safe: -> decompressed size, or -1 on a malformed stream / an output that does not fit
	inline int decompress_block(const std::uint8_t* src, int srcSize, std::uint8_t* dst, int cap)
	{
		if (srcSize <= 0 || cap < 0)
			return -1;
		if (cap == 0)
			return (srcSize == 1 && src[0] == 0) ? 0 : -1;
		const std::uint8_t* ip = src;
		const std::uint8_t* const iend = src + srcSize;
		std::uint8_t* op = dst;
		std::uint8_t* const oend = dst + cap;
		for (;;)
		{
			if (ip >= iend)
				return -1;
			const unsigned token = *ip++;
			// literals
			std::size_t length = token >> 4;
			if (length == 15)
			{
				unsigned s;
				do
				{
					if (ip >= iend)
						return -1;
					s = *ip++;
					length += s;
				} while (s == 255);
			}
			if (length > std::size_t(iend - ip) || length > std::size_t(oend - op))
				return -1;
			// literals that reach into the last 12 output bytes or the last 8 input bytes can only belong to the last sequence
			// (a match starts at least 12 bytes before the end and is followed by a token and 5 literals)
			const bool mustBeLast = length + MFLIMIT > std::size_t(oend - op) || length + 8 > std::size_t(iend - ip);
			if (mustBeLast && length != std::size_t(iend - ip))
				return -1;
			std::memcpy(op, ip, length);
			ip += length;
			op += length;
			if (mustBeLast)
				break; // the last sequence holds literals only
			const unsigned offset = ip[0] | (unsigned(ip[1]) << 8);
			ip += 2;
			if (offset == 0 || offset > std::size_t(op - dst))
				return -1;
			length = token & 15;
			if (length == 15)
			{
				unsigned s;
				do
				{
					if (ip >= iend)
						return -1;
					s = *ip++;
					length += s;
				} while (s == 255);
			}
			length += MINMATCH;
			if (length + LASTLITERALS > std::size_t(oend - op))
				return -1; // the last 5 bytes of the output must be literals
			const std::uint8_t* match = op - offset;
			for (std::size_t i = 0; i < length; ++i) // byte by byte: the match may overlap its own output
				op[i] = match[i];
			op += length;
		}
		return int(op - dst);
	}
}
