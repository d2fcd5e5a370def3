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

// ------------------------------------------------------------------ local alignment tables
// Letter codes of the local aligner (usb_local.cuh).  Every character class the substitution
// matrix (setnucmx.cpp:33-87 / blosum62.cpp:17-96) and the identity matrix (alpha2.cpp:220-279)
// distinguish gets one 6-bit code:
//   0..25 = 'A'..'Z', 26..51 = 'a'..'z', 52 = '-' or '.', 53 = '*', 54 = anything else.
// (Upper and lower case stay apart because the amino identity matrix sets B~N, B~D, Z~Q, Z~E for
// the upper-case characters only, alpha2.cpp:269-279.)
#define USB_NCODE 64
struct LocalTables {
	uint8_t code[256];
	int8_t score[USB_NCODE][USB_NCODE];  // substitution scores (integers, checked by the caller)
	uint64_t match[USB_NCODE];           // identity matrix, bit b of match[a]
	uint8_t word_letter[USB_NCODE];      // LocalAligner2 word letter; wildcards are letter 0 (localaligner2.cpp:95-97)
	uint8_t udb_letter[256];             // UDB word letter by raw character, 0xff = bad (udbparams.cpp:546-552)
	uint32_t alpha;                      // 4 or 20
};

// NCBI BLOSUM62 in 1/2-bit units, letters in the order of kBlosumOrder.
static const char kBlosumOrder[] = "ARNDCQEGHILKMFPSTWYVBZX*";
static const int8_t kBlosum62[24][24] = {
	{4, -1, -2, -2, 0, -1, -1, 0, -2, -1, -1, -1, -1, -2, -1, 1, 0, -3, -2, 0, -2, -1, 0, -4},
	{-1, 5, 0, -2, -3, 1, 0, -2, 0, -3, -2, 2, -1, -3, -2, -1, -1, -3, -2, -3, -1, 0, -1, -4},
	{-2, 0, 6, 1, -3, 0, 0, 0, 1, -3, -3, 0, -2, -3, -2, 1, 0, -4, -2, -3, 3, 0, -1, -4},
	{-2, -2, 1, 6, -3, 0, 2, -1, -1, -3, -4, -1, -3, -3, -1, 0, -1, -4, -3, -3, 4, 1, -1, -4},
	{0, -3, -3, -3, 9, -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1, -3, -3, -2, -4},
	{-1, 1, 0, 0, -3, 5, 2, -2, 0, -3, -2, 1, 0, -3, -1, 0, -1, -2, -1, -2, 0, 3, -1, -4},
	{-1, 0, 0, 2, -4, 2, 5, -2, 0, -3, -3, 1, -2, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4},
	{0, -2, 0, -1, -3, -2, -2, 6, -2, -4, -4, -2, -3, -3, -2, 0, -2, -2, -3, -3, -1, -2, -1, -4},
	{-2, 0, 1, -1, -3, 0, 0, -2, 8, -3, -3, -1, -2, -1, -2, -1, -2, -2, 2, -3, 0, 0, -1, -4},
	{-1, -3, -3, -3, -1, -3, -3, -4, -3, 4, 2, -3, 1, 0, -3, -2, -1, -3, -1, 3, -3, -3, -1, -4},
	{-1, -2, -3, -4, -1, -2, -3, -4, -3, 2, 4, -2, 2, 0, -3, -2, -1, -2, -1, 1, -4, -3, -1, -4},
	{-1, 2, 0, -1, -3, 1, 1, -2, -1, -3, -2, 5, -1, -3, -1, 0, -1, -3, -2, -2, 0, 1, -1, -4},
	{-1, -1, -2, -3, -1, 0, -2, -3, -2, 1, 2, -1, 5, 0, -2, -1, -1, -1, -1, 1, -3, -1, -1, -4},
	{-2, -3, -3, -3, -2, -3, -3, -3, -1, 0, 0, -3, 0, 6, -4, -2, -2, 1, 3, -1, -3, -3, -1, -4},
	{-1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7, -1, -1, -4, -3, -2, -2, -1, -2, -4},
	{1, -1, 1, 0, -1, 0, 0, 0, -1, -2, -2, 0, -1, -2, -1, 4, 1, -3, -2, -2, 0, 0, 0, -4},
	{0, -1, 0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1, 5, -2, -2, 0, -1, -1, 0, -4},
	{-3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1, -4, -3, -2, 11, 2, -3, -4, -3, -2, -4},
	{-2, -2, -2, -3, -2, -1, -2, -3, 2, -1, -1, -2, -1, 3, -3, -2, -2, 2, 7, -1, -3, -2, -1, -4},
	{0, -3, -3, -3, -1, -2, -2, -3, -3, 3, 1, -2, 1, -1, -2, -2, 0, -3, -1, 4, -3, -2, -1, -4},
	{-2, -1, 3, 4, -3, 0, 1, -1, 0, -3, -4, 0, -3, -3, -2, 0, -1, -4, -3, -3, 4, 1, -1, -4},
	{-1, 0, 0, 1, -3, 3, 4, -2, 0, -3, -3, 1, -1, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4},
	{0, -1, -1, -1, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -2, 0, 0, -2, -1, -1, -1, -1, -1, -4},
	{-4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, 1}};

inline uint32_t local_code_of(int c)
{
	if (c >= 'A' && c <= 'Z')
		return (uint32_t)(c - 'A');
	if (c >= 'a' && c <= 'z')
		return (uint32_t)(26 + c - 'a');
	if (c == '-' || c == '.')
		return 52;
	if (c == '*')
		return 53;
	return 54;
}

// match/mismatch: nucleotide substitution scores (ignored for amino acids, which use BLOSUM62).
inline void build_local_tables(bool nucleo, int match, int mismatch, LocalTables &L)
{
	CharTables T;
	build_char_tables(T);
	for (int c = 0; c < 256; ++c) {
		L.code[c] = (uint8_t)local_code_of(c);
		L.udb_letter[c] = 0xff;
	}
	for (int a = 0; a < USB_NCODE; ++a) {
		L.match[a] = 0;
		L.word_letter[a] = 0;
		for (int b = 0; b < USB_NCODE; ++b)
			L.score[a][b] = 0;
	}
	L.alpha = nucleo ? 4 : 20;
	// representative character of a code
	auto rep = [](int code) {
		return code < 26 ? 'A' + code : code < 52 ? 'a' + code - 26 : code == 52 ? '-' : code == 53 ? '*' : '?';
	};
	auto both_cases = [](char up, auto fn) {
		fn((int)local_code_of(up));
		fn((int)local_code_of(up + 32));
	};
	if (nucleo) {
		const char *nt = "ACGTU";
		for (int i = 0; i < 5; ++i) {
			const int li = i == 4 ? 3 : i;
			L.udb_letter[(int)nt[i]] = (uint8_t)li;
			both_cases(nt[i], [&](int ca) {
				L.word_letter[ca] = (uint8_t)li;
				for (int j = 0; j < 5; ++j)
					both_cases(nt[j], [&](int cb) { L.score[ca][cb] = (int8_t)(li == (j == 4 ? 3 : j) ? match : mismatch); });
			});
		}
		for (int a = 0; a < 55; ++a)
			for (int b = 0; b < 55; ++b)
				if (chars_match(T, (uint8_t)rep(a), (uint8_t)rep(b)))
					L.match[a] |= 1ull << b;
	} else {
		const char *aa = "ACDEFGHIKLMNPQRSTVWY"; // alpha.cpp g_CharToLetterAmino order
		for (int i = 0; i < 20; ++i) {
			L.udb_letter[(int)aa[i]] = (uint8_t)i;
			both_cases(aa[i], [&](int ca) { L.word_letter[ca] = (uint8_t)i; });
		}
		for (int i = 0; i < 24; ++i)
			for (int j = 0; j < 24; ++j) {
				const char x = kBlosumOrder[i], y = kBlosumOrder[j];
				const int v = kBlosum62[i][j];
				if (x == '*' || y == '*') {
					// '*' has no case; letters paired with it take both cases (blosum62.cpp:62-86)
					if (x == '*' && y == '*')
						L.score[53][53] = (int8_t)v;
					else if (x == '*')
						both_cases(y, [&](int cb) { L.score[53][cb] = (int8_t)v; L.score[cb][53] = (int8_t)v; });
					continue;
				}
				both_cases(x, [&](int ca) { both_cases(y, [&](int cb) { L.score[ca][cb] = (int8_t)v; }); });
			}
		// alpha2.cpp:250-279: letters match when equal ignoring case or when either is X/x;
		// B~N, B~D, Z~Q, Z~E for the upper-case characters only; non-letters only match as gap
		// symbol vs gap symbol
		for (int a = 0; a < 52; ++a)
			for (int b = 0; b < 52; ++b)
				if (a % 26 == b % 26 || a % 26 == 'X' - 'A' || b % 26 == 'X' - 'A')
					L.match[a] |= 1ull << b;
		auto pair = [&](char x, char y) {
			L.match[x - 'A'] |= 1ull << (y - 'A');
			L.match[y - 'A'] |= 1ull << (x - 'A');
		};
		pair('B', 'N'); pair('B', 'D'); pair('Z', 'Q'); pair('Z', 'E');
		L.match[52] |= 1ull << 52;
	}
}

} // namespace usb
