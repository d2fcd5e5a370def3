// usb_tables.h -- character tables shared by the host index builder and the CUDA kernels.
//
// Semantics follow the reference's alphabet tables (alpha.cpp g_CharToLetterNucleo /
// g_CharToCompChar, alpha2.cpp:94-150 IUPAC bit sets, alpha2.cpp:220-264 g_MatchMxNucleo);
// the tables themselves are rebuilt here from the IUPAC definitions.
#pragma once
#include <stdint.h>

namespace usb {

// Per-character class word:
//   bits 0..3  : single-base bit set (A=1 C=2 G=4 T/U=8), only for the unambiguous letters
//   bits 4..7  : IUPAC ambiguity set (superset of the above)
//   bit  8     : isalpha
//   bit  9     : gap character ('-' or '.')
struct CharTables {
	uint16_t cls[256];
	uint8_t upper[256];  // toupper for ASCII letters, identity otherwise
	uint8_t comp[256];   // complement; characters without one map to themselves (seqinfo.cpp:292-325)
};

inline void build_char_tables(CharTables &T)
{
	for (int c = 0; c < 256; ++c) {
		bool lo = (c >= 'a' && c <= 'z'), up = (c >= 'A' && c <= 'Z');
		T.cls[c] = (uint16_t)((lo || up) ? 0x100 : 0);
		T.upper[c] = (uint8_t)(lo ? c - 32 : c);
		T.comp[c] = (uint8_t)c;
	}
	T.cls[(int)'-'] |= 0x200;
	T.cls[(int)'.'] |= 0x200;
	auto single = [&](char ch, int b) {
		T.cls[(int)ch] |= (uint16_t)(b | (b << 4));
		T.cls[(int)ch + 32] |= (uint16_t)(b | (b << 4));
	};
	single('A', 1); single('C', 2); single('G', 4); single('T', 8); single('U', 8);
	auto base_bit = [](char ch) { return ch == 'A' ? 1 : ch == 'C' ? 2 : ch == 'G' ? 4 : 8; };
	const char *amb[] = {"MAC", "RAG", "WAT", "SCG", "YCT", "KGT", "VACG", "HACT", "DAGT", "BCGT", "XGATC", "NGATC"};
	for (const char *a : amb) {
		int set = 0;
		for (const char *p = a + 1; *p; ++p)
			set |= base_bit(*p);
		T.cls[(int)a[0]] |= (uint16_t)(set << 4);
		T.cls[(int)a[0] + 32] |= (uint16_t)(set << 4);
	}
	// complement pairs; 'u' (lower case) has no entry in the reference table and is kept as is
	const char *pairs = "ATBVCGDHGCHDKMMKNNRYSSTAUAVBWWXXYR";
	for (const char *p = pairs; *p; p += 2) {
		T.comp[(int)p[0]] = (uint8_t)p[1];
		if (p[0] != 'U')
			T.comp[(int)p[0] + 32] = (uint8_t)(p[1] + 32);
	}
}

// Identity test used for %id (alpha2.cpp:220-264): IUPAC-overlap counts as a match.
inline bool chars_match(const CharTables &T, uint8_t a, uint8_t b)
{
	uint16_t ca = T.cls[a], cb = T.cls[b];
	if (!(ca & 0x100) || !(cb & 0x100))
		return (ca & 0x200) && (cb & 0x200);
	if (T.upper[a] == T.upper[b])
		return true;
	return ((ca & 0xf) & (cb >> 4)) || ((cb & 0xf) & (ca >> 4));
}

} // namespace usb
