// TEST INFRASTRUCTURE ONLY (see simt_emu.h): the element-wise kernels of the library (one thread per element: fills, gathers, flags,
// emitters, score conversions) from their CUDA SOURCES under the SIMT emulator, each against the obvious loop, with arrays of exactly
// n elements and n not a multiple of the block -- so that, under build.py's sanitizers, an index past the end or a conflicting store
// is seen.  usage: elem_emu; exit code 0 = everything identical
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "cuda_runtime.h"  // tests/simt/shim -> simt_emu.h

#include "xxm_kernels.cuh"
#include "xxm_sweep.cuh"
#include "xxm_stage.cuh"  // k_scatter_stage
#include "xxm_xchg.cuh"   // k_pack_hist

namespace xxm {
alignas(16) uint8_t smemM[1 << 16];
uint32_t s_hs[1 << 14];
}  // namespace xxm

using namespace xxm;

namespace {
int failures = 0;
void check(const char* what, bool ok) {
	std::printf("%-22s %s\n", what, ok ? "identical" : "DIFFERENT");
	failures += ok ? 0 : 1;
}
template <class F>
void each(uint32_t n, F&& body) { emu::launch((n + 255) / 256, 256, body); }
}  // namespace

int main() {
	std::mt19937 rng(101);
	const uint32_t n = 1237, nval = 77;
	std::uniform_real_distribution<double> U(0.0, 1.0);
	std::vector<uint32_t> seq(n), pos(n), vidx(n), idx(n);
	std::vector<float> score(n);
	std::vector<uint8_t> keep(n);
	for (uint32_t i = 0; i < n; ++i) {
		seq[i] = 1 + i / 5; pos[i] = 1 + rng() % 300; vidx[i] = (rng() % 9 == 0) ? kNone : rng() % nval;
		score[i] = static_cast<float>(U(rng)); keep[i] = rng() % 3 != 0; idx[i] = rng() % n;
	}
	std::vector<double> pS(nval), fin(nval), posProb(400), perCand(n);
	for (uint32_t v = 0; v < nval; ++v) { pS[v] = (v % 7 == 0) ? 1.0 : U(rng) * 1e-2; fin[v] = (v % 5 == 0) ? 1.0 : U(rng); }
	for (double& v : posProb) v = 0.1 + 0.9 * U(rng);
	for (double& v : perCand) v = U(rng);
	{  // fills and iota
		std::vector<uint32_t> a(n, 1u), c(n, 1u);
		std::vector<float> b(n, 1.f);
		each(n, [&] { k_fill_u32(a.data(), 0xC0FFEEu, n); });
		each(n, [&] { k_fill_f32(b.data(), 2.5f, n); });
		each(n, [&] { k_iota(c.data(), n); });
		bool ok = true;
		for (uint32_t i = 0; i < n; ++i) ok = ok && a[i] == 0xC0FFEEu && b[i] == 2.5f && c[i] == i;
		check("k_fill_u32/f32, k_iota", ok);
	}
	{  // gathers
		std::vector<uint32_t> g(n), keys(n), vals(n);
		std::vector<double> src(n), gd(n);
		for (uint32_t i = 0; i < n; ++i) src[i] = U(rng);
		each(n, [&] { k_gather_u32(pos.data(), idx.data(), n, g.data()); });
		each(n, [&] { k_gather_f64(src.data(), idx.data(), n, gd.data()); });
		each(n, [&] { k_gather_keys(score.data(), idx.data(), n, keys.data(), vals.data()); });
		bool ok = true;
		for (uint32_t i = 0; i < n; ++i) {
			uint32_t bits;
			std::memcpy(&bits, &score[idx[i]], 4);
			ok = ok && g[i] == pos[idx[i]] && gd[i] == src[idx[i]] && keys[i] == bits && vals[i] == idx[i];
		}
		check("k_gather_u32/f64/keys", ok);
	}
	{  // flags by score
		bool ok = true;
		for (int strict = 0; strict < 2; ++strict)
			for (int withKeep = 0; withKeep < 2; ++withKeep) {
				std::vector<uint8_t> f(n, 0xEE);
				const float thr = score[17];
				each(n, [&] { k_flag_score(score.data(), withKeep ? keep.data() : nullptr, n, thr, strict, f.data()); });
				for (uint32_t i = 0; i < n; ++i) ok = ok && f[i] == ((withKeep ? keep[i] != 0 : true) && (strict ? score[i] < thr : score[i] <= thr));
			}
		check("k_flag_score", ok);
	}
	for (int kind = 0; kind < 3; ++kind) {  // values of list entries under the three kinds of finalPval
		Values V{};
		V.pS = pS.data(); V.finalTab = fin.data(); V.posProb = posProb.data(); V.perCand = perCand.data();
		V.kind = kind == 0 ? kFinalTable : kind == 1 ? kFinalLocMops : kFinalPerCand;
		const int shift = kind == 1 ? -1 : 0;
		std::vector<uint8_t> f(n, 0xEE);
		std::vector<float> sP(n, -1.f), sF(n, -1.f);
		std::vector<double> dF(n, -1.0), dP(n, -1.0);
		uint32_t count = 0;
		each(n, [&] { k_flag_hits(V, vidx.data(), pos.data(), n, f.data(), &count); });
		each(n, [&] { k_entry_scores(V, vidx.data(), pos.data(), n, shift, 0, sP.data()); });
		each(n, [&] { k_entry_scores(V, vidx.data(), pos.data(), n, shift, 1, sF.data()); });
		each(n, [&] { k_entry_doubles(V, vidx.data(), pos.data(), n, dF.data(), dP.data()); });
		bool ok = true;
		for (uint32_t i = 0; i < n; ++i) {
			const uint32_t v = vidx[i];
			const double f0 = (v == kNone) ? 1.0 : final_of(V, v, i, static_cast<int>(pos[i]));
			const double fs = (v == kNone) ? 1.0 : final_of(V, v, i, static_cast<int>(pos[i]) + shift);
			const double p = (v == kNone) ? 1.0 : pS[v];
			const double fd = (v == kNone) ? 1.0 : ((p == 1.0) ? 1.0 : f0);
			ok = ok && f[i] == ((v != kNone) && f0 < 1.0) && sP[i] == static_cast<float>(p) && sF[i] == static_cast<float>(fs) && dF[i] == fd && dP[i] == p;
		}
		check(kind == 0 ? "values, table" : kind == 1 ? "values, localized" : "values, per candidate", ok);
	}
	{  // site bits + flags through them
		std::vector<uint32_t> bits((nval + 31) / 32, 0xEEEEEEEEu);
		std::vector<uint8_t> f(n, 0xEE);
		each(nval, [&] { k_site_bits(fin.data(), nval, bits.data()); });
		each(n, [&] { k_flag_hits_bits(bits.data(), vidx.data(), n, f.data()); });
		bool ok = true;
		for (uint32_t v = 0; v < nval; ++v) ok = ok && (((bits[v >> 5] >> (v & 31)) & 1u) == (fin[v] < 1.0 ? 1u : 0u));
		for (uint32_t i = 0; i < n; ++i) ok = ok && f[i] == ((vidx[i] != kNone) && fin[vidx[i]] < 1.0);
		check("k_site_bits/flag_hits_bits", ok);
	}
	{  // histogram over value indices of kept entries
		std::vector<uint32_t> v2(vidx), h(nval, 0u), want(nval, 0u);
		for (uint32_t i = 0; i < n; ++i) { if (v2[i] == kNone) v2[i] = 0; if (keep[i]) ++want[v2[i]]; }
		each(n, [&] { k_hist_vidx(v2.data(), keep.data(), n, h.data()); });
		check("k_hist_vidx", h == want);
	}
	{  // memo bookkeeping
		const uint32_t cap = 1000;
		std::vector<int32_t> key(cap);
		std::vector<uint32_t> state(cap), sidx;
		for (uint32_t h = 0; h < cap; ++h) { key[h] = (rng() % 3 == 0) ? kMemoEmpty : static_cast<int32_t>(rng()); state[h] = rng() % 4; }
		bool ok = true;
		for (uint32_t bit = 0; bit < 2; ++bit) {
			std::vector<uint8_t> f(cap, 0xEE);
			each(cap, [&] { k_memo_flag(key.data(), state.data(), cap, bit, f.data()); });
			for (uint32_t h = 0; h < cap; ++h) ok = ok && f[h] == ((key[h] != kMemoEmpty) && !((state[h] >> bit) & 1u));
		}
		for (uint32_t h = 0; h < cap; h += 3) sidx.push_back(h);  // distinct slots
		std::vector<double> vals(sidx.size()), dst(cap, -1.0), wantD(cap, -1.0);
		std::vector<uint32_t> wantS(state);
		for (size_t k = 0; k < sidx.size(); ++k) { vals[k] = U(rng); wantD[sidx[k]] = vals[k]; wantS[sidx[k]] |= 2u; }
		const uint32_t m = static_cast<uint32_t>(sidx.size());
		each(m, [&] { k_scatter_mode(vals.data(), sidx.data(), m, dst.data(), state.data()); });
		check("k_memo_flag, k_scatter_mode", ok && dst == wantD && state == wantS);
	}
	{  // site records
		std::vector<SiteRec> out(n), out2(n);
		each(n, [&] { k_emit_sites(seq.data(), pos.data(), score.data(), idx.data(), n, 3, out.data()); });
		each(n, [&] { k_emit_sites(seq.data(), pos.data(), score.data(), nullptr, n, -2, out2.data()); });
		bool ok = true;
		for (uint32_t i = 0; i < n; ++i) {
			const uint32_t e = idx[i];
			ok = ok && out[i].score == score[e] && out[i].startPos == static_cast<int32_t>(pos[e]) + 3 && out[i].sequence == static_cast<int32_t>(seq[e]);
			ok = ok && out2[i].score == score[i] && out2[i].startPos == static_cast<int32_t>(pos[i]) - 2 && out2[i].sequence == static_cast<int32_t>(seq[i]);
		}
		const uint32_t nslots = 311;
		std::vector<uint32_t> slotIdx(nslots), sel;
		std::vector<float> slotScore(nslots);
		for (uint32_t s = 0; s < nslots; ++s) { slotIdx[s] = (rng() % 4 == 0) ? kNone : rng() % n; slotScore[s] = slotIdx[s] == kNone ? 1.f : score[slotIdx[s]]; if (rng() % 2) sel.push_back(s); }
		std::vector<SiteRec> o3(nslots), o4(sel.size());
		each(nslots, [&] { k_emit_slot_sites(slotIdx.data(), slotScore.data(), pos.data(), 1, nullptr, nslots, o3.data()); });
		const uint32_t ns = static_cast<uint32_t>(sel.size());
		each(ns, [&] { k_emit_slot_sites(slotIdx.data(), slotScore.data(), pos.data(), 1, sel.data(), ns, o4.data()); });
		auto expect = [&](uint32_t s, const SiteRec& r) {
			const uint32_t e = slotIdx[s];
			return r.score == slotScore[s] && r.startPos == ((e == kNone) ? 1 : static_cast<int32_t>(pos[e]) + 1) && r.sequence == static_cast<int32_t>(s) + 1;
		};
		for (uint32_t s = 0; s < nslots; ++s) ok = ok && expect(s, o3[s]);
		for (uint32_t k = 0; k < ns; ++k) ok = ok && expect(sel[k], o4[k]);
		check("k_emit_sites/slot_sites", ok);
	}
	{  // dense rank table + membership bitmap from the compact (id, class) lists
		const int len = 5, nid = 300;
		const uint32_t size = 1u << (2 * len);
		std::vector<uint32_t> ids, cls;
		std::vector<uint8_t> clsOk(20);
		for (auto& o : clsOk) o = rng() % 4 != 0;
		std::vector<uint32_t> perm(size);
		for (uint32_t i = 0; i < size; ++i) perm[i] = i;
		std::shuffle(perm.begin(), perm.end(), rng);
		for (int i = 0; i < nid; ++i) { ids.push_back(perm[static_cast<size_t>(i)]); cls.push_back(rng() % 20); }
		std::vector<uint32_t> rank(size, 0u), bm(size / 32, 0u), wantR(size, 0u), wantB(size / 32, 0u);
		for (int i = 0; i < nid; ++i)
			if (clsOk[cls[static_cast<size_t>(i)]]) { wantR[ids[static_cast<size_t>(i)]] = cls[static_cast<size_t>(i)] + 1; wantB[ids[static_cast<size_t>(i)] >> 5] |= 1u << (ids[static_cast<size_t>(i)] & 31); }
		each(static_cast<uint32_t>(nid), [&] { k_fill_rank(ids.data(), cls.data(), clsOk.data(), nid, rank.data(), bm.data()); });
		check("k_fill_rank", rank == wantR && bm == wantB);
	}
	{  // the scatter half of the batched table staging: segments of a staging buffer to their arrays (word copies where both sides are
		// aligned, byte copies otherwise), blockIdx.y = segment
		std::vector<uint8_t> stage(5000);
		for (auto& b : stage) b = static_cast<uint8_t>(rng());
		const uint32_t off[4] = {0, 1024, 2051, 4000}, bytes[4] = {1000, 1026, 1500, 1};  // aligned words, aligned with a tail, unaligned source, one byte
		std::vector<std::vector<uint32_t>> dst(4);
		UpSegs segs{};
		for (int k = 0; k < 4; ++k) {
			dst[k].assign((bytes[k] + 3) / 4 + 1, 0xEEEEEEEEu);
			segs.dst[k] = (k == 3) ? static_cast<void*>(reinterpret_cast<uint8_t*>(dst[k].data()) + 1) : static_cast<void*>(dst[k].data());  // ... and an unaligned destination
			segs.off[k] = off[k]; segs.bytes[k] = bytes[k];
		}
		for (int k = 0; k < 4; ++k) {  // (the emulator's grid is one-dimensional: one launch per segment, blockIdx.y = k)
			emu::blockY = static_cast<unsigned>(k);
			emu::launch(2, 256, [&] { k_scatter_stage(stage.data(), segs); });
		}
		emu::blockY = 0;
		bool ok = true;
		for (int k = 0; k < 4; ++k) {
			const uint8_t* d = static_cast<const uint8_t*>(segs.dst[k]);
			ok = ok && !std::memcmp(d, stage.data() + off[k], bytes[k]) && d[bytes[k]] == 0xEE;  // and not one byte more
		}
		check("k_scatter_stage", ok);
	}
	{  // the exchange vector of a sharded step: header (two slots taken from the device) + histogram padded to its fixed length
		bool ok = true;
		for (uint32_t nv : {1u, 300u, kXchgHist}) {
			std::vector<uint32_t> hist(nv);
			for (auto& h : hist) h = rng() % 100000;
			XchgHead head{};
			for (uint32_t i = 0; i < kXchgHead; ++i) head.v[i] = 1000 + i;
			const uint32_t kept = 4242;
			const unsigned long long tot = 9000, cap = 8000;
			for (int fromDev = 0; fromDev < 2; ++fromDev) {
				std::vector<unsigned long long> out(kXchgHead + kXchgHist, ~0ull);
				emu::launch((kXchgHist + 255) / 256, 256, [&] { k_pack_hist(hist.data(), nv, head, fromDev ? &kept : nullptr, fromDev ? &tot : nullptr, cap, out.data()); });
				for (uint32_t i = 0; i < kXchgHead; ++i) {
					unsigned long long w = head.v[i];
					if (fromDev && i == 5) w = kept;
					if (fromDev && i == 8) w = 1;  // the list outgrew its arrays
					ok = ok && out[i] == w;
				}
				for (uint32_t i = 0; i < kXchgHist; ++i) ok = ok && out[kXchgHead + i] == (i < nv ? hist[i] : 0ull);
			}
		}
		check("k_pack_hist", ok);
	}
	std::printf("%d failed\n", failures);
	return failures ? 1 : 0;
}
