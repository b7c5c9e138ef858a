/*
 * bt_native.cuh — rank queries on the UNMODIFIED .ebwt side layout, and the re-layout of that
 * layout into the kernel's 32-byte rank blocks.
 *
 * Native layout (reference ebwt.h:164-167, 4203-4281; consumers 2034-2226): 64-byte sides of
 * 56 B 2-bit BWT + 2 x u32 occ; even sides are "backward" (filled from the last byte to the
 * first, high bit-pair first) and carry occ[A], occ[C]; odd sides are "forward" and carry occ[G],
 * occ[T]; all four occ words of a pair are cumulative counts at the pair's midpoint.
 *
 * Used once per index load (relayout kernel); the search kernels only see BtBlock.
 */
#pragma once
#include "bt_core.cuh"

struct BtNativeIndex {
	const uint8_t *ebwt;
	uint32_t len, zOff, zEbwtByteOff, zEbwtBpOff;
	uint32_t fchr[5];
};

/* SideLocus::initFromRow (ebwt.h:1469-1497) */
struct BtNativeLocus { uint32_t sideByteOff; uint32_t fw, by, bp; };
BT_FN BtNativeLocus bt_native_locus(uint32_t row) {
	BtNativeLocus l;
	uint32_t sideNum = row / 224u, charOff = row % 224u;
	l.sideByteOff = sideNum * 64u;
	l.fw = sideNum & 1u;
	l.by = charOff >> 2; l.bp = charOff & 3u;
	if (!l.fw) { l.by = 55u - l.by; l.bp ^= 3u; }
	return l;
}
/* Ebwt::rowL */
BT_FN uint32_t bt_native_row_l(const BtNativeIndex &ix, uint32_t row) {
	BtNativeLocus l = bt_native_locus(row);
	return (ix.ebwt[l.sideByteOff + l.by] >> (2 * l.bp)) & 3u;
}
/* countUpToEx + countFwSideEx / countBwSideEx (ebwt.h:1963-2027, 2081-2129, 2184-2226):
 * out[c] = fchr[c] + #c in BWT[0,row), '$' excluded — i.e. one half of mapLFEx. */
BT_FN void bt_native_lf_ex(const BtNativeIndex &ix, uint32_t row, uint32_t out[4]) {
	BtNativeLocus l = bt_native_locus(row);
	const uint8_t *side = ix.ebwt + l.sideByteOff;
	uint32_t cnt[4] = { 0, 0, 0, 0 };
	/* whole 8-byte words below byte `by`, then whole bytes, then the low `bp` bit-pairs of byte `by` */
	uint32_t i = 0;
	for (; i + 7 < l.by; i += 8) {
		uint64_t w = 0;
		for (uint32_t k = 0; k < 8; k++) w |= (uint64_t)side[i + k] << (8 * k);
		uint64_t lo = w & 0x5555555555555555ull, hi = (w >> 1) & 0x5555555555555555ull;
		uint32_t t = (uint32_t)BT_POPC64(hi & lo), g = (uint32_t)BT_POPC64(hi) - t, c = (uint32_t)BT_POPC64(lo) - t;
		cnt[3] += t; cnt[2] += g; cnt[1] += c; cnt[0] += 32 - t - g - c;
	}
	for (; i < l.by; i++) { uint32_t b = side[i]; cnt[b & 3]++; cnt[(b >> 2) & 3]++; cnt[(b >> 4) & 3]++; cnt[(b >> 6) & 3]++; }
	for (uint32_t k = 0; k < l.bp; k++) cnt[(side[l.by] >> (2 * k)) & 3]++;
	uint32_t B = l.sideByteOff + l.by;
	bool zHere = (l.sideByteOff <= ix.zEbwtByteOff) && (B >= ix.zEbwtByteOff);
	if (l.fw) {
		if (zHere && (B > ix.zEbwtByteOff || (B == ix.zEbwtByteOff && l.bp > ix.zEbwtBpOff))) cnt[0]--;
		const uint8_t *ac = side - 8, *gt = side + 56;
		uint32_t occ[4];
		for (uint32_t k = 0; k < 2; k++) {
			occ[k] = (uint32_t)ac[4 * k] | ((uint32_t)ac[4 * k + 1] << 8) | ((uint32_t)ac[4 * k + 2] << 16) | ((uint32_t)ac[4 * k + 3] << 24);
			occ[2 + k] = (uint32_t)gt[4 * k] | ((uint32_t)gt[4 * k + 1] << 8) | ((uint32_t)gt[4 * k + 2] << 16) | ((uint32_t)gt[4 * k + 3] << 24);
		}
		for (uint32_t c = 0; c < 4; c++) out[c] = occ[c] + cnt[c] + ix.fchr[c];
	} else {
		cnt[(side[l.by] >> (2 * l.bp)) & 3]++;
		if (zHere && (B > ix.zEbwtByteOff || (B == ix.zEbwtByteOff && l.bp >= ix.zEbwtBpOff))) cnt[0]--;
		const uint8_t *ac = side + 56, *gt = side + 120;
		uint32_t occ[4];
		for (uint32_t k = 0; k < 2; k++) {
			occ[k] = (uint32_t)ac[4 * k] | ((uint32_t)ac[4 * k + 1] << 8) | ((uint32_t)ac[4 * k + 2] << 16) | ((uint32_t)ac[4 * k + 3] << 24);
			occ[2 + k] = (uint32_t)gt[4 * k] | ((uint32_t)gt[4 * k + 1] << 8) | ((uint32_t)gt[4 * k + 2] << 16) | ((uint32_t)gt[4 * k + 3] << 24);
		}
		for (uint32_t c = 0; c < 4; c++) out[c] = occ[c] - cnt[c] + ix.fchr[c];
	}
}

/* Build rank block k (rows 64k .. 64k+63) from the native layout. */
BT_FN void bt_relayout_block(const BtNativeIndex &ix, uint32_t k, uint4 out[2]) {
	uint32_t row0 = k << 6;
	uint32_t occ[4];
	bt_native_lf_ex(ix, row0, occ);
	uint64_t hi = 0, lo = 0;
	for (uint32_t o = 0; o < 64; o++) {
		uint32_t row = row0 + o;
		if (row > ix.len) break;                     /* rows past the '$' row are never queried */
		uint32_t c = bt_native_row_l(ix, row);
		hi |= (uint64_t)(c >> 1) << o;
		lo |= (uint64_t)(c & 1) << o;
	}
	out[0].x = occ[0]; out[0].y = occ[1]; out[0].z = occ[2]; out[0].w = occ[3];
	out[1].x = (uint32_t)hi; out[1].y = (uint32_t)(hi >> 32); out[1].z = (uint32_t)lo; out[1].w = (uint32_t)(lo >> 32);
}
