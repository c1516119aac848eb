// Host-side check of the HDLC tables of gnuais_b200/csrc/gais_track.cuh (no GPU needed): stepping the
// FSM four bits at a time through hdlc_nibble_entry() -- with the field layout the kernel reads (row
// offset, k in byte 2, stored bits in byte 3, ENTER/EMIT flags, closing-bit position) -- must do exactly
// what stepping it bit by bit through hdlc_transition() does: same next state, same stored bits in the
// same order, same frame starts and ends.  Exit code 0 = ok.
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include "gais_track.cuh"
#include "gais_oracle.h"

using namespace gais;

struct Trace {
	uint32_t id;
	std::vector<int> stored;      // every stored bit, in order
	std::vector<long> enters, emits;   // bit indices
};

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t) (rng_state >> 32); }

int main()
{
	// a bit stream with everything in it: idle noise, long alternations (training), flags, stuffed payloads
	std::vector<int> bits;
	for (int burst = 0; burst < 4000; burst++) {
		int idle = rnd() % 40;
		for (int i = 0; i < idle; i++) bits.push_back(rnd() & 1);
		int train = 16 + rnd() % 16;
		for (int i = 0; i < train; i++) bits.push_back(i & 1);
		const int flag[8] = { 0, 1, 1, 1, 1, 1, 1, 0 };
		for (int i = 0; i < 8; i++) bits.push_back(flag[i]);
		int ones = 0;
		if (burst % 3 == 0) {
			// a well-formed frame: 21 or 53 payload bytes + CRC-16/X.25 (low byte first), LSB first, bit-stuffed
			uint8_t bytes[55];
			const int nb = (burst % 6 == 0) ? 53 : 21;
			for (int j = 0; j < nb; j++) bytes[j] = (uint8_t) rnd();
			const uint16_t fcs = goracle_crc16(bytes, (unsigned) nb);
			bytes[nb] = (uint8_t) (fcs & 0xff);
			bytes[nb + 1] = (uint8_t) (fcs >> 8);
			for (int j = 0; j < nb + 2; j++)
				for (int k = 0; k < 8; k++) {
					const int b = (bytes[j] >> k) & 1;
					bits.push_back(b);
					ones = b ? ones + 1 : 0;
					if (ones == 5) { bits.push_back(0); ones = 0; }
				}
			for (int i = 0; i < 8; i++) bits.push_back(flag[i]);
			continue;
		}
		int len = rnd() % 470;
		for (int i = 0; i < len; i++) {
			int b = (rnd() % 3) != 0;            // biased to ones: exercises the stuffing paths
			bits.push_back(b);
			ones = b ? ones + 1 : 0;
			if (ones == 5 && (rnd() % 8)) { bits.push_back(0); ones = 0; }   // stuff (usually)
		}
		if (rnd() % 4) for (int i = 0; i < 8; i++) bits.push_back(flag[i]);
	}
	while (bits.size() % 4) bits.push_back(0);

	for (uint32_t start = 0; start < (uint32_t) H_NSTATES; start++) {
		Trace a, b;
		a.id = b.id = start;
		// (a) bit by bit
		for (size_t i = 0; i < bits.size(); i++) {
			const uint32_t e = hdlc_transition(a.id, (uint32_t) bits[i]);
			a.id = e & 0x7fu;
			if (e & H_STORE) a.stored.push_back(bits[i]);
			if (e & H_ENTER) { a.enters.push_back((long) i); }
			if (e & H_EMIT) a.emits.push_back((long) i);
		}
		// (b) four bits at a time, reading the entry the way track_kernel does
		uint32_t row = b.id << 6;
		for (size_t i = 0; i < bits.size(); i += 4) {
			const uint32_t v = (uint32_t) (bits[i] | bits[i + 1] << 1 | bits[i + 2] << 2 | bits[i + 3] << 3);
			const uint32_t e = hdlc_nibble_entry(row >> 6, v);
			const uint32_t k = (e >> 16) & 0xffu, st = e >> 24;
			if (k > 4 || (e & 0x3cu) || ((e >> 19) & 0x1fu)) { printf("bad entry %08x\n", e); return 1; }
			if (e & N_ENTER) b.enters.push_back(-1);         // position inside the nibble is not recorded for ENTER
			for (uint32_t j = 0; j < k; j++) b.stored.push_back((int) ((st >> j) & 1u));
			if (e & N_EMIT) b.emits.push_back((long) (i + (e & 3u)));
			row = e & N_ROW;
		}
		b.id = row >> 6;
		if (a.id != b.id || a.emits != b.emits || a.enters.size() != b.enters.size()) {
			printf("start state %u: id %u/%u, %zu/%zu emits, %zu/%zu enters\n", start, a.id, b.id, a.emits.size(), b.emits.size(),
			       a.enters.size(), b.enters.size());
			return 1;
		}
		// the stored bits differ only in that an ENTER inside a nibble discards what the nibble stored before it
		// (the frame buffer restarts): compare the bits stored after the LAST enter, frame by frame, via the totals
		if (a.stored.size() < b.stored.size()) { printf("start state %u: nibble path stored more bits\n", start); return 1; }
	}
	// frame-exact comparison from the reset state: rebuild every frame's bits on both paths
	{
		std::vector<std::vector<int>> fa, fb;
		std::vector<int> cur;
		uint32_t id = hdlc_hunt_id(0, 0, 0);
		for (size_t i = 0; i < bits.size(); i++) {
			const uint32_t e = hdlc_transition(id, (uint32_t) bits[i]);
			id = e & 0x7fu;
			if (e & H_ENTER) cur.clear();
			if (e & H_STORE) cur.push_back(bits[i]);
			if (e & H_EMIT) { fa.push_back(cur); cur.clear(); }
		}
		cur.clear();
		uint32_t row = hdlc_hunt_id(0, 0, 0) << 6;
		for (size_t i = 0; i < bits.size(); i += 4) {
			const uint32_t v = (uint32_t) (bits[i] | bits[i + 1] << 1 | bits[i + 2] << 2 | bits[i + 3] << 3);
			const uint32_t e = hdlc_nibble_entry(row >> 6, v);
			if (e & N_ENTER) cur.clear();
			for (uint32_t j = 0; j < ((e >> 16) & 0xffu); j++) cur.push_back((int) ((e >> (24 + j)) & 1u));
			if (e & N_EMIT) { fb.push_back(cur); cur.clear(); }
			row = e & N_ROW;
		}
		if (fa.size() != fb.size() || fa.size() < 1000) { printf("frames %zu/%zu\n", fa.size(), fb.size()); return 1; }
		for (size_t f = 0; f < fa.size(); f++)
			if (fa[f] != fb[f]) { printf("frame %zu differs (%zu / %zu bits)\n", f, fa[f].size(), fb[f].size()); return 1; }
		printf("ok: %zu bits, %zu frames, %d start states\n", bits.size(), fa.size(), H_NSTATES);
	}
	// against the oracle's bit machine (oracle/gais_oracle.c fsm_bit(), pinned to the reference): the table
	// FSM with the kernel's buffer rules (position reset on ENTER / EMIT, reset at 449 stored bits, a closed
	// frame judged by stop bit, length and CRC) must report the same frame events
	{
		std::vector<uint8_t> b8(bits.begin(), bits.end());
		std::vector<goracle_frame> want(100000);
		int32_t stats[3];
		int64_t n_want = 0;
		if (goracle_fsm_bits(b8.data(), (int64_t) b8.size(), stats, want.data(), (int64_t) want.size(), &n_want) != 0) return 1;
		uint32_t id = hdlc_hunt_id(0, 0, 0), pos = 0;
		uint8_t store[450] = { 0 };
		int64_t n_got = 0;
		int32_t got_stats[3] = { 0, 0, 0 };
		for (size_t i = 0; i < bits.size(); i++) {
			const uint32_t bit = (uint32_t) bits[i], e = hdlc_transition(id, bit);
			id = e & 0x7fu;
			if (e & H_STORE) {                                   // as hdlc_bits_serial() in gais_track.cuh
				store[pos++] = (uint8_t) bit;
				if (pos >= 449u) { id = hdlc_hunt_id(0, 0, bit); pos = 0; }
			}
			if (e & H_ENTER) { pos = 0; for (int k = 0; k < 450; k++) store[k] = 0; }
			if (e & H_EMIT) {
				const int nbits = (int) pos - 22;
				int status = 2;
				if (bit == 0 && nbits > 0) {                     // as frame_check_kernel
					uint8_t bytes[60];
					const int nb = nbits / 8;
					for (int j = 0; j < nb + 2; j++) {
						unsigned v = 0;
						for (int k = 0; k < 8; k++) v |= (unsigned) store[8 * j + k] << k;
						bytes[j] = (uint8_t) v;
					}
					status = goracle_crc16(bytes, (unsigned) nb + 2) == 0x0f47 ? 0 : 1;
				}
				got_stats[status]++;
				if (n_got < n_want) {
					const goracle_frame &w = want[n_got];
					if (w.end_bit != (uint32_t) i || w.nbits != (int16_t) nbits || w.status != status) {
						printf("frame %lld: end %u/%zu nbits %d/%d status %d/%d\n", (long long) n_got, w.end_bit, i, w.nbits, nbits, w.status,
						       status);
						return 1;
					}
				}
				n_got++;
				pos = 0;
			}
		}
		if (n_got != n_want || got_stats[0] != stats[0] || got_stats[1] != stats[1] || got_stats[2] != stats[2] || n_got < 1000) {
			printf("oracle: %lld frames (%d/%d/%d), tables: %lld (%d/%d/%d)\n", (long long) n_want, stats[0], stats[1], stats[2],
			       (long long) n_got, got_stats[0], got_stats[1], got_stats[2]);
			return 1;
		}
		printf("ok: %lld frame events equal to the oracle's (%d crc-ok, %d crc-fail, %d size-fail)\n", (long long) n_got, stats[0], stats[1],
		       stats[2]);
	}
	return 0;
}
