// mbjson.h -- a small JSON value, parser and writer for the host mirror and the CLI.
// (The reference uses nlohmann::json; this is an independent minimal implementation.  Objects
// keep their keys sorted, as nlohmann's default std::map-backed objects do, so round-tripped
// metadata prints in the same key order as `boss`.)
#ifndef MB_HOST_JSON_H
#define MB_HOST_JSON_H

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace MachineBoss {

struct Json {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  double num = 0;
  bool isInt = false;
  std::string str;
  std::vector<Json> arr;
  std::map<std::string, Json> obj;

  Json() {}
  static Json string (const std::string& s) { Json j; j.type = String; j.str = s; return j; }
  static Json number (double d) { Json j; j.type = Number; j.num = d; j.isInt = (d == std::floor (d) && std::fabs (d) < 1e15); return j; }

  bool isNull() const { return type == Null; }
  bool has (const std::string& k) const { return type == Object && obj.count (k); }
  const Json& at (const std::string& k) const {
    if (type != Object || !obj.count (k)) throw std::runtime_error ("JSON: missing key \"" + k + "\"");
    return obj.at (k);
  }
  const Json& at (size_t n) const {
    if (type != Array || n >= arr.size()) throw std::runtime_error ("JSON: array index out of range");
    return arr[n];
  }
  size_t size() const { return type == Array ? arr.size() : type == Object ? obj.size() : 0; }
  const std::string& asString() const { if (type != String) throw std::runtime_error ("JSON: expected a string"); return str; }
  double asNumber() const {
    if (type == Number) return num;
    if (type == String) {   // the reference's infinity-safe strings (jsonio.h:14-22)
      if (str == "-Infinity") return -INFINITY;
      if (str == "Infinity") return INFINITY;
      if (str == "NaN") return NAN;
    }
    throw std::runtime_error ("JSON: expected a number");
  }
  long long asInt() const { return (long long) asNumber(); }

  static std::string escape (const std::string& s) {
    std::string o;
    for (unsigned char c: s) {
      switch (c) {
        case '"': o += "\\\""; break;
        case '\\': o += "\\\\"; break;
        case '\n': o += "\\n"; break;
        case '\t': o += "\\t"; break;
        case '\r': o += "\\r"; break;
        default:
          if (c < 0x20) { char buf[8]; snprintf (buf, sizeof buf, "\\u%04x", c); o += buf; }
          else o += (char) c;
      }
    }
    return o;
  }

  void write (std::ostream& out) const {
    switch (type) {
      case Null: out << "null"; break;
      case Bool: out << (b ? "true" : "false"); break;
      case Number:
        if (isInt) out << (long long) num;
        else { char buf[40]; snprintf (buf, sizeof buf, "%.17g", num); out << buf; }
        break;
      case String: out << '"' << escape (str) << '"'; break;
      case Array: {
        out << '[';
        for (size_t n = 0; n < arr.size(); ++n) { if (n) out << ','; arr[n].write (out); }
        out << ']';
        break;
      }
      case Object: {
        out << '{';
        size_t n = 0;
        for (const auto& kv: obj) { if (n++) out << ','; out << '"' << escape (kv.first) << "\":"; kv.second.write (out); }
        out << '}';
        break;
      }
    }
  }
  std::string dump() const { std::ostringstream o; write (o); return o.str(); }

  // ---- parser ----
  static Json parse (const std::string& text) {
    size_t p = 0;
    Json j = parseValue (text, p);
    skipWs (text, p);
    if (p != text.size()) throw std::runtime_error ("JSON: trailing characters at offset " + std::to_string (p));
    return j;
  }

private:
  static void skipWs (const std::string& t, size_t& p) { while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\t' || t[p] == '\r')) ++p; }
  static void fail (const std::string& what, size_t p) { throw std::runtime_error ("JSON: " + what + " at offset " + std::to_string (p)); }
  static Json parseValue (const std::string& t, size_t& p) {
    skipWs (t, p);
    if (p >= t.size()) fail ("unexpected end", p);
    Json j;
    const char c = t[p];
    if (c == '{') {
      j.type = Object;
      ++p; skipWs (t, p);
      if (p < t.size() && t[p] == '}') { ++p; return j; }
      for (;;) {
        skipWs (t, p);
        if (p >= t.size() || t[p] != '"') fail ("expected a key", p);
        const std::string k = parseString (t, p);
        skipWs (t, p);
        if (p >= t.size() || t[p] != ':') fail ("expected ':'", p);
        ++p;
        j.obj[k] = parseValue (t, p);
        skipWs (t, p);
        if (p < t.size() && t[p] == ',') { ++p; continue; }
        if (p < t.size() && t[p] == '}') { ++p; break; }
        fail ("expected ',' or '}'", p);
      }
    } else if (c == '[') {
      j.type = Array;
      ++p; skipWs (t, p);
      if (p < t.size() && t[p] == ']') { ++p; return j; }
      for (;;) {
        j.arr.push_back (parseValue (t, p));
        skipWs (t, p);
        if (p < t.size() && t[p] == ',') { ++p; continue; }
        if (p < t.size() && t[p] == ']') { ++p; break; }
        fail ("expected ',' or ']'", p);
      }
    } else if (c == '"') {
      j.type = String;
      j.str = parseString (t, p);
    } else if (t.compare (p, 4, "true") == 0) { j.type = Bool; j.b = true; p += 4; }
    else if (t.compare (p, 5, "false") == 0) { j.type = Bool; j.b = false; p += 5; }
    else if (t.compare (p, 4, "null") == 0) { p += 4; }
    else {
      const char* s = t.c_str() + p;
      char* e = nullptr;
      const double d = strtod (s, &e);
      if (e == s) fail ("unexpected character", p);
      j.type = Number;
      j.num = d;
      j.isInt = true;
      for (const char* q = s; q < e; ++q) if (*q == '.' || *q == 'e' || *q == 'E') j.isInt = false;
      p += (size_t) (e - s);
    }
    return j;
  }
  static std::string parseString (const std::string& t, size_t& p) {
    std::string o;
    ++p;
    while (p < t.size() && t[p] != '"') {
      if (t[p] == '\\') {
        ++p;
        if (p >= t.size()) fail ("bad escape", p);
        switch (t[p]) {
          case 'n': o += '\n'; break;
          case 't': o += '\t'; break;
          case 'r': o += '\r'; break;
          case 'b': o += '\b'; break;
          case 'f': o += '\f'; break;
          case 'u': {
            if (p + 4 >= t.size()) fail ("bad \\u escape", p);
            const unsigned cp = (unsigned) strtoul (t.substr (p + 1, 4).c_str(), nullptr, 16);
            if (cp < 0x80) o += (char) cp;
            else if (cp < 0x800) { o += (char) (0xC0 | (cp >> 6)); o += (char) (0x80 | (cp & 0x3F)); }
            else { o += (char) (0xE0 | (cp >> 12)); o += (char) (0x80 | ((cp >> 6) & 0x3F)); o += (char) (0x80 | (cp & 0x3F)); }
            p += 4;
            break;
          }
          default: o += t[p];
        }
        ++p;
      } else o += t[p++];
    }
    if (p >= t.size()) fail ("unterminated string", p);
    ++p;
    return o;
  }
};

}  // namespace MachineBoss

#endif
