// msgpack_mini.h — MessagePack encoder / decoder for the Json value of json_mini.h and a gzip wrapper: the container of the
// reference's snapshots (Testbed::save_snapshot, src/testbed.cu:5288-5355: nlohmann::json::to_msgpack, wrapped in zstr's gzip
// stream for ".ingp").  Encoding choices follow nlohmann's writer (smallest integer type; float32 when the value is exactly
// representable, else float64; objects in key order) so files are byte-comparable for the same content.
#pragma once

#include <zlib.h>

#include <cstring>

#include "json_mini.h"

namespace ngpb {

class MsgPackWriter {
public:
	std::vector<uint8_t> out;
	void write(const Json& j) {
		switch (j.type) {
			case Json::Null: out.push_back(0xC0); break;
			case Json::Bool: out.push_back(j.b ? 0xC3 : 0xC2); break;
			case Json::Number: number(j); break;
			case Json::String: {
				const size_t n = j.str.size();
				if (n <= 31) out.push_back((uint8_t)(0xA0 | n));
				else if (n <= 0xFF) { out.push_back(0xD9); be(n, 1); }
				else if (n <= 0xFFFF) { out.push_back(0xDA); be(n, 2); }
				else { out.push_back(0xDB); be(n, 4); }
				out.insert(out.end(), j.str.begin(), j.str.end());
			} break;
			case Json::Array: {
				const size_t n = j.arr.size();
				if (n <= 15) out.push_back((uint8_t)(0x90 | n));
				else if (n <= 0xFFFF) { out.push_back(0xDC); be(n, 2); }
				else { out.push_back(0xDD); be(n, 4); }
				for (const auto& v : j.arr) write(v);
			} break;
			case Json::Object: {
				const size_t n = j.obj.size();
				if (n <= 15) out.push_back((uint8_t)(0x80 | n));
				else if (n <= 0xFFFF) { out.push_back(0xDE); be(n, 2); }
				else { out.push_back(0xDF); be(n, 4); }
				for (const auto& kv : j.obj) {
					Json k;
					k.type = Json::String;
					k.str = kv.first;
					write(k);
					write(kv.second);
				}
			} break;
			case Json::Binary: {
				const size_t n = j.bin.size();
				if (n <= 0xFF) { out.push_back(0xC4); be(n, 1); }
				else if (n <= 0xFFFF) { out.push_back(0xC5); be(n, 2); }
				else { out.push_back(0xC6); be(n, 4); }
				out.insert(out.end(), j.bin.begin(), j.bin.end());
			} break;
		}
	}

private:
	void be(uint64_t v, int bytes) {
		for (int i = bytes - 1; i >= 0; --i) out.push_back((uint8_t)(v >> (8 * i)));
	}
	void number(const Json& j) {
		if (j.integer) {
			if (j.num >= 0) {
				const uint64_t v = (uint64_t)j.num;
				if (v <= 0x7F) out.push_back((uint8_t)v);
				else if (v <= 0xFF) { out.push_back(0xCC); be(v, 1); }
				else if (v <= 0xFFFF) { out.push_back(0xCD); be(v, 2); }
				else if (v <= 0xFFFFFFFFull) { out.push_back(0xCE); be(v, 4); }
				else { out.push_back(0xCF); be(v, 8); }
			} else {
				const int64_t v = (int64_t)j.num;
				if (v >= -32) out.push_back((uint8_t)(int8_t)v);
				else if (v >= -128) { out.push_back(0xD0); be((uint64_t)v, 1); }
				else if (v >= -32768) { out.push_back(0xD1); be((uint64_t)v, 2); }
				else if (v >= -2147483648ll) { out.push_back(0xD2); be((uint64_t)v, 4); }
				else { out.push_back(0xD3); be((uint64_t)v, 8); }
			}
			return;
		}
		const float f = (float)j.num;
		if ((double)f == j.num) {
			uint32_t u;
			memcpy(&u, &f, 4);
			out.push_back(0xCA);
			be(u, 4);
		} else {
			uint64_t u;
			memcpy(&u, &j.num, 8);
			out.push_back(0xCB);
			be(u, 8);
		}
	}
};

class MsgPackReader {
public:
	MsgPackReader(const uint8_t* data, size_t n) : d(data), n(n) {}
	// nesting is bounded: a crafted file must not be able to exhaust the stack (not catchable) through recursion
	struct DepthGuard {
		int& d;
		explicit DepthGuard(int& depth) : d(depth) {
			if (++d > 64) throw std::runtime_error("msgpack: nesting deeper than 64 levels");
		}
		~DepthGuard() { --d; }
	};
	int depth = 0;
	Json read() {
		DepthGuard guard(depth);
		const uint8_t c = byte();
		Json j;
		if (c <= 0x7F) return jint(c);
		if (c >= 0xE0) return jint((int8_t)c);
		if ((c & 0xF0) == 0x80) return map(c & 0x0F);
		if ((c & 0xF0) == 0x90) return array(c & 0x0F);
		if ((c & 0xE0) == 0xA0) return string(c & 0x1F);
		switch (c) {
			case 0xC0: return Json{};
			case 0xC2: return jbool(false);
			case 0xC3: return jbool(true);
			case 0xC4: return binary(be(1));
			case 0xC5: return binary(be(2));
			case 0xC6: return binary(be(4));
			case 0xCA: { const uint32_t u = (uint32_t)be(4); float f; memcpy(&f, &u, 4); return jnum(f); }
			case 0xCB: { const uint64_t u = be(8); double f; memcpy(&f, &u, 8); return jnum(f); }
			case 0xCC: return jint((int64_t)be(1));
			case 0xCD: return jint((int64_t)be(2));
			case 0xCE: return jint((int64_t)be(4));
			case 0xCF: { Json v = jint(0); v.num = (double)be(8); return v; }
			case 0xD0: return jint((int8_t)be(1));
			case 0xD1: return jint((int16_t)be(2));
			case 0xD2: return jint((int32_t)be(4));
			case 0xD3: return jint((int64_t)be(8));
			case 0xD9: return string(be(1));
			case 0xDA: return string(be(2));
			case 0xDB: return string(be(4));
			case 0xDC: return array(be(2));
			case 0xDD: return array(be(4));
			case 0xDE: return map(be(2));
			case 0xDF: return map(be(4));
			// ext family (nlohmann binary with a subtype): keep the payload as binary
			case 0xD4: byte(); return binary(1);
			case 0xD5: byte(); return binary(2);
			case 0xD6: byte(); return binary(4);
			case 0xD7: byte(); return binary(8);
			case 0xD8: byte(); return binary(16);
			case 0xC7: { const size_t len = be(1); byte(); return binary(len); }
			case 0xC8: { const size_t len = be(2); byte(); return binary(len); }
			case 0xC9: { const size_t len = be(4); byte(); return binary(len); }
			default: throw std::runtime_error("msgpack: unsupported type byte");
		}
	}
	bool at_end() const { return p == n; }

private:
	const uint8_t* d;
	size_t n, p = 0;
	uint8_t byte() {
		if (p >= n) throw std::runtime_error("msgpack: truncated");
		return d[p++];
	}
	uint64_t be(int bytes) {
		uint64_t v = 0;
		for (int i = 0; i < bytes; ++i) v = (v << 8) | byte();
		return v;
	}
	void need(size_t len) {
		if (len > n - p) throw std::runtime_error("msgpack: truncated");
	}
	Json string(size_t len) {
		need(len);
		Json j = jstr(std::string((const char*)d + p, len));
		p += len;
		return j;
	}
	Json binary(size_t len) {
		need(len);
		Json j = jbin(d + p, len);
		p += len;
		return j;
	}
	Json array(size_t len) {
		Json j = jarr();
		if (len > n - p) throw std::runtime_error("msgpack: truncated input (array longer than the remaining bytes)");   // every element takes at least one byte
		j.arr.reserve(len);
		for (size_t i = 0; i < len; ++i) j.arr.push_back(read());
		return j;
	}
	Json map(size_t len) {
		Json j = jobj();
		if (len > (n - p) / 2) throw std::runtime_error("msgpack: truncated input (map longer than the remaining bytes)");
		for (size_t i = 0; i < len; ++i) {
			Json k = read();
			if (k.type != Json::String) throw std::runtime_error("msgpack: non-string map key");
			j.obj[k.str] = read();
		}
		return j;
	}
};

// gzip (deflate with a gzip header, window bits 15+16: what zstr::ostream writes) / gunzip with zlib-gzip auto-detection
// (15+32: what zstr::istream reads)
inline std::vector<uint8_t> gzip_compress(const std::vector<uint8_t>& in, int level = Z_DEFAULT_COMPRESSION) {
	z_stream zs{};
	if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("zlib: deflateInit2 failed");
	std::vector<uint8_t> out(deflateBound(&zs, (uLong)in.size()) + 64);
	zs.next_in = const_cast<Bytef*>(in.data());
	zs.avail_in = (uInt)in.size();
	zs.next_out = out.data();
	zs.avail_out = (uInt)out.size();
	const int rc = deflate(&zs, Z_FINISH);
	deflateEnd(&zs);
	if (rc != Z_STREAM_END) throw std::runtime_error("zlib: deflate failed");
	out.resize(zs.total_out);
	return out;
}
inline std::vector<uint8_t> gzip_decompress(const std::vector<uint8_t>& in) {
	z_stream zs{};
	if (inflateInit2(&zs, 15 + 32) != Z_OK) throw std::runtime_error("zlib: inflateInit2 failed");
	std::vector<uint8_t> out(in.size() * 4 + 1024);
	zs.next_in = const_cast<Bytef*>(in.data());
	zs.avail_in = (uInt)in.size();
	size_t produced = 0;
	for (;;) {
		zs.next_out = out.data() + produced;
		zs.avail_out = (uInt)(out.size() - produced);
		const int rc = inflate(&zs, Z_NO_FLUSH);
		produced = out.size() - zs.avail_out;
		if (rc == Z_STREAM_END) break;
		if (rc != Z_OK && rc != Z_BUF_ERROR) {
			inflateEnd(&zs);
			throw std::runtime_error("zlib: inflate failed (not a gzip / zlib stream?)");
		}
		if (zs.avail_out == 0) out.resize(out.size() * 2);
		else if (zs.avail_in == 0) { inflateEnd(&zs); throw std::runtime_error("zlib: truncated stream"); }
	}
	inflateEnd(&zs);
	out.resize(produced);
	return out;
}

// JSON text of a value (binary values as {"bytes": n}); for diagnostics and the codec tests
inline void json_dump(const Json& j, std::string& o) {
	switch (j.type) {
		case Json::Null: o += "null"; break;
		case Json::Bool: o += j.b ? "true" : "false"; break;
		case Json::Number: {
			char buf[40];
			if (j.integer) snprintf(buf, sizeof(buf), "%lld", (long long)j.num);
			else snprintf(buf, sizeof(buf), "%.17g", j.num);
			o += buf;
		} break;
		case Json::String: {
			o += '"';
			for (char c : j.str) {
				if (c == '"' || c == '\\') { o += '\\'; o += c; }
				else if (c == '\n') o += "\\n";
				else if ((unsigned char)c < 0x20u) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", (unsigned)(unsigned char)c); o += b; }
				else o += c;
			}
			o += '"';
		} break;
		case Json::Array: {
			o += '[';
			for (size_t i = 0; i < j.arr.size(); ++i) {
				if (i) o += ',';
				json_dump(j.arr[i], o);
			}
			o += ']';
		} break;
		case Json::Object: {
			o += '{';
			bool first = true;
			for (const auto& kv : j.obj) {
				if (!first) o += ',';
				first = false;
				json_dump(jstr(kv.first), o);
				o += ':';
				json_dump(kv.second, o);
			}
			o += '}';
		} break;
		case Json::Binary: o += "{\"bytes\":" + std::to_string(j.bin.size()) + "}"; break;
	}
}

}  // namespace ngpb
