// json_mini.h — a small recursive-descent JSON reader for network configs (configs/nerf/base.json style).
// The reference parses these with nlohmann::json and tolerates comments (src/testbed.cu:304); so does this.
#pragma once

#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ngpb {

struct Json {
	enum Type { Null, Bool, Number, String, Array, Object, Binary } type = Null;
	bool b = false;
	double num = 0.0;
	bool integer = false;           // Number written without fraction / exponent (msgpack distinguishes integers from floats)
	std::string str;
	std::vector<uint8_t> bin;       // Binary (msgpack bin; nlohmann::json::binary_t in the reference's snapshots)
	std::vector<Json> arr;
	std::map<std::string, Json> obj;

	bool contains(const std::string& k) const { return type == Object && obj.count(k) > 0; }
	const Json& at(const std::string& k) const {
		auto it = obj.find(k);
		if (type != Object || it == obj.end()) throw std::runtime_error("json: missing key '" + k + "'");
		return it->second;
	}
	double value(const std::string& k, double def) const { return contains(k) && obj.at(k).type == Number ? obj.at(k).num : def; }
	std::string value(const std::string& k, const std::string& def) const { return contains(k) && obj.at(k).type == String ? obj.at(k).str : def; }
	const Json& sub(const std::string& k) const {
		static const Json empty{};
		return contains(k) ? obj.at(k) : empty;
	}
};

class JsonParser {
public:
	explicit JsonParser(const std::string& text) : s(text) {}
	Json parse() {
		Json v = value();
		skip();
		if (p != s.size()) fail("trailing characters");
		return v;
	}

private:
	const std::string& s;
	size_t p = 0;
	[[noreturn]] void fail(const std::string& m) { throw std::runtime_error("json parse error at byte " + std::to_string(p) + ": " + m); }
	void skip() {
		for (;;) {
			while (p < s.size() && std::isspace((unsigned char)s[p])) ++p;
			if (p + 1 < s.size() && s[p] == '/' && s[p + 1] == '/') {
				while (p < s.size() && s[p] != '\n') ++p;
			} else if (p + 1 < s.size() && s[p] == '/' && s[p + 1] == '*') {
				p += 2;
				while (p + 1 < s.size() && !(s[p] == '*' && s[p + 1] == '/')) ++p;
				p += 2;
			} else {
				break;
			}
		}
	}
	Json value() {
		skip();
		if (p >= s.size()) fail("unexpected end");
		const char c = s[p];
		if (c == '{') return object();
		if (c == '[') return array();
		if (c == '"') {
			Json j;
			j.type = Json::String;
			j.str = string();
			return j;
		}
		if (s.compare(p, 4, "true") == 0) { p += 4; Json j; j.type = Json::Bool; j.b = true; return j; }
		if (s.compare(p, 5, "false") == 0) { p += 5; Json j; j.type = Json::Bool; j.b = false; return j; }
		if (s.compare(p, 4, "null") == 0) { p += 4; return Json{}; }
		char* end = nullptr;
		const size_t p0 = p;
		const double d = std::strtod(s.c_str() + p, &end);
		if (end == s.c_str() + p) fail("unexpected token");
		p = (size_t)(end - s.c_str());
		Json j;
		j.type = Json::Number;
		j.num = d;
		j.integer = true;
		for (const char* q = s.c_str() + p0; q < end; ++q)
			if (*q == '.' || *q == 'e' || *q == 'E' || *q == 'n' || *q == 'i') j.integer = false;
		return j;
	}
	uint32_t hex4() {  // p is on the 'u' of a \uXXXX escape; leaves p on its last hex digit
		if (p + 4 >= s.size()) fail("truncated \\u escape");
		uint32_t v = 0;
		for (int k = 1; k <= 4; ++k) {
			const char ch = s[p + k];
			v <<= 4;
			if (ch >= '0' && ch <= '9') v |= (uint32_t)(ch - '0');
			else if (ch >= 'a' && ch <= 'f') v |= (uint32_t)(ch - 'a' + 10);
			else if (ch >= 'A' && ch <= 'F') v |= (uint32_t)(ch - 'A' + 10);
			else fail("bad hex digit in \\u escape");
		}
		p += 4;
		return v;
	}
	std::string string() {
		++p;
		std::string out;
		while (p < s.size() && s[p] != '"') {
			if (s[p] == '\\' && p + 1 < s.size()) {
				++p;
				switch (s[p]) {
					case 'n': out += '\n'; break;
					case 't': out += '\t'; break;
					case 'r': out += '\r'; break;
					case 'b': out += '\b'; break;
					case 'f': out += '\f'; break;
					case 'u': {
						uint32_t cp = hex4();
						if (cp >= 0xD800u && cp < 0xDC00u && p + 2 < s.size() && s[p + 1] == '\\' && s[p + 2] == 'u') {  // surrogate pair
							p += 2;
							const uint32_t lo = hex4();
							if (lo < 0xDC00u || lo > 0xDFFFu) fail("bad low surrogate in \\u escape");
							cp = 0x10000u + ((cp - 0xD800u) << 10) + (lo - 0xDC00u);
						}
						if (cp < 0x80u) {
							out += (char)cp;
						} else if (cp < 0x800u) {
							out += (char)(0xC0u | (cp >> 6));
							out += (char)(0x80u | (cp & 0x3Fu));
						} else if (cp < 0x10000u) {
							out += (char)(0xE0u | (cp >> 12));
							out += (char)(0x80u | ((cp >> 6) & 0x3Fu));
							out += (char)(0x80u | (cp & 0x3Fu));
						} else {
							out += (char)(0xF0u | (cp >> 18));
							out += (char)(0x80u | ((cp >> 12) & 0x3Fu));
							out += (char)(0x80u | ((cp >> 6) & 0x3Fu));
							out += (char)(0x80u | (cp & 0x3Fu));
						}
						break;
					}
					default: out += s[p]; break;
				}
			} else {
				out += s[p];
			}
			++p;
		}
		if (p >= s.size()) fail("unterminated string");
		++p;
		return out;
	}
	Json array() {
		Json j;
		j.type = Json::Array;
		++p;
		skip();
		if (p < s.size() && s[p] == ']') { ++p; return j; }
		for (;;) {
			j.arr.push_back(value());
			skip();
			if (p < s.size() && s[p] == ',') { ++p; continue; }
			if (p < s.size() && s[p] == ']') { ++p; break; }
			fail("expected , or ]");
		}
		return j;
	}
	Json object() {
		Json j;
		j.type = Json::Object;
		++p;
		skip();
		if (p < s.size() && s[p] == '}') { ++p; return j; }
		for (;;) {
			skip();
			if (p >= s.size() || s[p] != '"') fail("expected key");
			const std::string k = string();
			skip();
			if (p >= s.size() || s[p] != ':') fail("expected :");
			++p;
			j.obj[k] = value();
			skip();
			if (p < s.size() && s[p] == ',') { ++p; continue; }
			if (p < s.size() && s[p] == '}') { ++p; break; }
			fail("expected , or }");
		}
		return j;
	}
};

// ---- builders ---------------------------------------------------------------------------------------------------------
inline Json jnum(double v) { Json j; j.type = Json::Number; j.num = v; return j; }
inline Json jint(int64_t v) { Json j; j.type = Json::Number; j.num = (double)v; j.integer = true; return j; }
inline Json jbool(bool v) { Json j; j.type = Json::Bool; j.b = v; return j; }
inline Json jstr(const std::string& v) { Json j; j.type = Json::String; j.str = v; return j; }
inline Json jobj() { Json j; j.type = Json::Object; return j; }
inline Json jarr() { Json j; j.type = Json::Array; return j; }
inline Json jbin(const void* data, size_t n) {
	Json j;
	j.type = Json::Binary;
	j.bin.assign((const uint8_t*)data, (const uint8_t*)data + n);
	return j;
}
inline Json jvec(const float* v, int n) {
	Json j = jarr();
	for (int i = 0; i < n; ++i) j.arr.push_back(jnum(v[i]));
	return j;
}

inline std::string to_lower(std::string v) {
	for (auto& c : v) c = (char)std::tolower((unsigned char)c);
	return v;
}

}  // namespace ngpb
