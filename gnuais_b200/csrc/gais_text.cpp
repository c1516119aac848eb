/*
 * gais_text.cpp -- host-side restatement of the per-message text line gnuais prints to stdout
 * (SURVEY.md 8f row N2), fed from gais_msg records instead of struct demod_state_t:
 *
 *     ch %c type %d mmsi %09ld: <fields of the message type> (!AIVDM,...)\n
 *
 * Reference: protodec_getdata() src/protodec.c:931-985 (header, dispatch, trailer) and the
 * field decoders protodec_pos/4/5/6/7_13/8/18/19/20/24 + DAC 1 FI 11/40, src/protodec.c:216-776.
 * Only the stdout text is reproduced; the MySQL / cache / range sinks those functions also
 * feed are out of scope.  Quirks kept on purpose: fields are
 * read from the zero-filled bit buffer even beyond the message length; `rateofturn`/`navstat`
 * are plain (signed) char; the FI 11 decoder's offsets (msg_start += 16 after a 25-bit field)
 * are the reference's own; type 19 prints two spaces before "width".
 */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gais_b200.h"
#include "gais_nmea.h"

namespace {

struct Bits {
	const uint8_t *p;
	int nbytes;
	/* protodec_henten(), src/protodec.c:205-214, over rbuffer as src/protodec.c:150-162 leaves it */
	unsigned long get(int from, int size) const
	{
		unsigned long v = 0;
		for (int i = 0; i < size; i++) {
			int k = from + i;
			unsigned long bit = (k < 8 * nbytes) ? (p[k >> 3] >> (7 - (k & 7))) & 1u : 0u;
			v |= bit << (size - 1 - i);
		}
		return v;
	}
};

/* protodec_decode_sixbit_ascii(), src/protodec.c:190-203 */
char sixbit_ascii(int c)
{
	if (c >= 1 && c <= 31)
		return (char) (c + 64);
	if (c >= 32 && c <= 63)
		return (char) c;
	return ' ';
}

/* n six-bit characters from bit `pos`, then remove_trailing_spaces(), src/protodec.c:173-184 */
void sixbit_string(const Bits &b, int pos, int n, char *out)
{
	for (int k = 0; k < n; k++, pos += 6)
		out[k] = sixbit_ascii((int) b.get(pos, 6));
	out[n] = 0;
	for (int i = n - 1; i >= 0 && (out[i] == ' ' || out[i] == 0); i--)
		out[i] = 0;
}

const char *appid_ifm(int i)          /* src/protodec.c:220-271 */
{
	switch (i) {
	case 0: return "text-telegram";
	case 1: return "application-ack";
	case 2: return "iai-fi-capab-interrogation";
	case 3: return "iai-capabi-interrogation";
	case 4: return "capability-reply";
	case 11: return "tide-weather";
	case 16: return "vts-targets";
	case 17: return "ship-waypoints";
	case 18: return "advice-of-waypoints";
	case 19: return "extended-ship-data";
	case 20: return "berthing-data";
	case 21: return "weather-obs-report";
	case 22: return "area-notice-bc";
	case 23: return "area-notice-addr";
	case 24: return "extended-ship-static";
	case 25: return "dangerous-cargo-info";
	case 26: return "environmental";
	case 27: return "route-info-bc";
	case 28: return "route-info-addr";
	case 29: return "text-description-bc";
	case 30: return "text-description-addr";
	case 40: return "persons-on-board";
	default: return "unknown";
	}
}

int sext(unsigned long v, int bits)    /* the "|= 0xF0000000" sign extensions of the reference */
{
	int x = (int) v;
	if ((x >> (bits - 1)) & 1)
		x |= (int) (~0u << bits);
	return x;
}

struct Out {
	char *s;
	int cap, n;
	void put(const char *fmt, ...) __attribute__((format(printf, 2, 3)));
};

void Out::put(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	if (n < cap) {
		int w = vsnprintf(s + n, (size_t) (cap - n), fmt, ap);
		n += (w < cap - n) ? w : cap - n - 1;
	}
	va_end(ap);
}

void msg_bin(Out &o, const Bits &b, int fi, int start)      /* src/protodec.c:277-342 */
{
	if (fi == 40) {
		o.put(" persons-on-board %d", (int) b.get(start, 13));
	} else if (fi == 11) {
		int ms = start;
		int latitude = (int) b.get(ms, 24);
		int longitude = (int) b.get(ms += 24, 25);
		int wind_speed = (int) b.get(ms += 16, 7);
		int wind_gust = (int) b.get(ms += 7, 7);
		int wind_dir = (int) b.get(ms += 7, 9);
		int wind_gust_dir = (int) b.get(ms += 9, 9);
		int air_temp = (int) b.get(ms += 9, 11);
		int rel_humid = (int) b.get(ms += 11, 7);
		int dew_point = (int) b.get(ms += 7, 10);
		int air_press = (int) b.get(ms += 10, 9) + 800;
		int air_press_tend = (int) b.get(ms += 9, 2);
		int horiz_visib_nm = (int) b.get(ms += 2, 8);
		int water_level = (int) b.get(ms += 8, 9);
		int wave_height = (int) b.get(ms += 5, 8);
		int water_temp = (int) b.get(ms += 4, 10);
		o.put(" lat %.6f lon %.6f wind_speed %dkt wind_gust %dkt wind_dir %d wind_gust_dir %d air_temp %.1fC rel_humid %d%% "
		      "dew_point %.1fC pressure %d pressure_tend %d visib %.1fNM water_level %.1fm wave_height %.1fm water_temp %.1fC",
		      (float) latitude / 60000.0, (float) longitude / 60000.0, wind_speed, wind_gust, wind_dir, wind_gust_dir,
		      (float) air_temp / 10.0 - 60.0, rel_humid, (float) dew_point / 10.0 - 20.0, air_press, air_press_tend,
		      (float) horiz_visib_nm / 10.0, (float) water_level / 10.0 - 10.0, (float) wave_height / 10.0,
		      (float) water_temp / 10.0 - 10.0);
	}
}

void position_line(Out &o, int latitude, int longitude, unsigned short course, unsigned short sog, char rateofturn, char navstat,
		   unsigned short heading)
{
	o.put(" lat %.6f lon %.6f course %.0f speed %.1f rateofturn %d navstat %d heading %d", (float) latitude / 600000.0,
	      (float) longitude / 600000.0, (float) course / 10.0, (float) sog / 10.0, rateofturn, navstat, heading);
}

} /* namespace */

extern "C" int gais_text_format(const gais_msg *m, char chanid, char *out, int cap)
{
	if (!m || !out || cap < 64)
		return GAIS_EINVAL;
	const int nbytes = m->nbits >> 3;
	Bits b = { m->payload, nbytes };
	const int type = (int) b.get(0, 6);
	if (type < 1 || type > 24)                         /* src/protodec.c:898-900 */
		return 0;
	const unsigned long mmsi = b.get(8, 30);
	const int fill = (m->nbits % 6) ? 6 - m->nbits % 6 : 0;
	const int bufferlen = m->nbits + fill;             /* src/protodec.c:909-915 */

	Out o = { out, cap, 0 };
	o.put("ch %c type %d mmsi %09ld:", chanid, type, (long) mmsi);   /* :934 */

	switch (type) {
	case 1: case 2: case 3: {                          /* protodec_pos, :349-395 */
		int longitude = sext(b.get(61, 28), 28), latitude = sext(b.get(38 + 22 + 29, 27), 27);
		position_line(o, latitude, longitude, (unsigned short) b.get(38 + 22 + 28 + 28, 12), (unsigned short) b.get(50, 10),
			      (char) b.get(38 + 2, 8), (char) b.get(38, 2), (unsigned short) b.get(38 + 22 + 28 + 28 + 12, 9));
		break;
	}
	case 4: {                                          /* protodec_4, :397-442 */
		unsigned long year = b.get(40, 12), month = b.get(52, 4), day = b.get(56, 5), hour = b.get(61, 5),
			      minute = b.get(66, 6), second = b.get(72, 6);
		int longitude = sext(b.get(79, 28), 28), latitude = sext(b.get(107, 27), 27);
		float longit = ((float) longitude) / 10000.0 / 60.0, latit = ((float) latitude) / 10000.0 / 60.0;
		o.put(" date %ld-%ld-%ld time %02ld:%02ld:%02ld lat %.6f lon %.6f", (long) year, (long) month, (long) day, (long) hour,
		      (long) minute, (long) second, latit, longit);
		break;
	}
	case 5: {                                          /* protodec_5, :444-521 */
		char name[21], destination[21];
		sixbit_string(b, 112, 20, name);
		sixbit_string(b, 120 + 106 + 68 + 8, 20, destination);
		unsigned int shiptype = (unsigned int) b.get(232, 8);
		unsigned int A = (unsigned int) b.get(240, 9), B = (unsigned int) b.get(249, 9);
		unsigned char C = (unsigned char) b.get(258, 6), D = (unsigned char) b.get(264, 6), draught = (unsigned char) b.get(294, 8);
		o.put(" name \"%s\" destination \"%s\" type %d length %d width %d draught %.1f", name, destination, shiptype, A + B, C + D,
		      (float) draught / 10.0);
		break;
	}
	case 6: {                                          /* protodec_6, :527-543 */
		int sequence = (int) b.get(38, 2);
		unsigned long dst = b.get(40, 30);
		int retransmitted = (int) b.get(70, 1), appid = (int) b.get(72, 16), dac = (int) b.get(72, 10), fi = (int) b.get(82, 6);
		o.put(" dst_mmsi %09ld seq %d retransmitted %d appid %d app_dac %d app_fi %d", (long) dst, sequence, retransmitted, appid,
		      dac, fi);
		if (dac == 1) {
			o.put("(%s)", appid_ifm(fi));
			msg_bin(o, b, fi, 88);
		}
		break;
	}
	case 7: case 13: {                                 /* protodec_7_13, :550-568 */
		int pos = 40;
		o.put(" buflen %d pos+32 %d", bufferlen, pos + 32);
		for (int i = 0; i < 4 && pos + 32 <= bufferlen; pos += 32) {
			o.put(" ack %d (to %09ld seq %d)", i + 1, (long) b.get(pos, 30), (int) b.get(pos + 30, 2));
			i++;
		}
		break;
	}
	case 8: {                                          /* protodec_8, :574-585 */
		int appid = (int) b.get(40, 16), dac = (int) b.get(40, 10), fi = (int) b.get(50, 6);
		o.put(" appid %d app_dac %d app_fi %d", appid, dac, fi);
		if (dac == 1) {
			o.put("(%s)", appid_ifm(fi));
			msg_bin(o, b, fi, 56);
		}
		break;
	}
	case 18: {                                         /* protodec_18, :587-632: rateofturn 0, navstat 15 */
		int longitude = sext(b.get(57, 28), 28), latitude = sext(b.get(85, 27), 27);
		position_line(o, latitude, longitude, (unsigned short) b.get(112, 12), (unsigned short) b.get(46, 10), 0, 15,
			      (unsigned short) b.get(124, 9));
		break;
	}
	case 19: {                                         /* protodec_19, :634-682 */
		char name[21];
		sixbit_string(b, 143, 20, name);
		unsigned int shiptype = (unsigned int) b.get(263, 8), A = (unsigned int) b.get(271, 9), B = (unsigned int) b.get(280, 9);
		unsigned char C = (unsigned char) b.get(289, 6), D = (unsigned char) b.get(295, 6);
		o.put(" name \"%s\" type %d length %d  width %d", name, shiptype, A + B, C + D);
		break;
	}
	case 20: {                                         /* protodec_20, :684-702 */
		int pos = 40;
		for (int i = 0; i < 4 && pos + 30 < bufferlen; pos += 30) {
			o.put(" reserve %d (ofs %d slots %d timeout %d incr %d)", i + 1, (int) b.get(pos, 12), (int) b.get(pos + 12, 4),
			      (int) b.get(pos + 16, 3), (int) b.get(pos + 19, 11));
			i++;
		}
		break;
	}
	case 24: {                                         /* protodec_24, :704-776 */
		int partnr = (int) b.get(38, 2);
		if (partnr == 0) {
			char name[21];
			sixbit_string(b, 40, 20, name);
			o.put(" name \"%s\"", name);
		}
		if (partnr == 1) {
			char callsign[7];
			sixbit_string(b, 90, 6, callsign);
			unsigned int shiptype = (unsigned int) b.get(40, 8), A = (unsigned int) b.get(132, 9), B = (unsigned int) b.get(141, 9);
			unsigned char C = (unsigned char) b.get(150, 6), D = (unsigned char) b.get(156, 6);
			o.put(" callsign \"%s\" type %d length %d width %d", callsign, shiptype, A + B, C + D);
		}
		break;
	}
	default:
		break;
	}

	/* trailer " (!%s)\n" with d->nmea = the LAST sentence protodec_generate_nmea() built (:984) */
	char text[GAIS_NMEA_STRIDE];
	int len = gn_format(m->payload, m->nbits, m->flags & 15, text);
	int start = 0;
	for (int i = 0; i + 2 < len; i++)
		if (text[i] == '\n')
			start = i + 1;
	text[len - 2] = 0;                                 /* drop "\r\n" */
	o.put(" (%s)\n", text + start);
	return o.n;
}
