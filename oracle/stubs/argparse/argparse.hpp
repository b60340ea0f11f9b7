// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// Minimal stand-in for p-ranav/argparse v3.2, which the reference FetchContent's
// from the network (reference CMakeLists.txt:16-20) and which is absent offline.
// It implements only the call surface that reference cmd/libtorch_bench/main.cpp:139-193
// uses, so that file can be compiled UNMODIFIED from where it lies under /root/reference
// (see oracle/Makefile). Written from the call sites, not from argparse's sources.
#pragma once
#include <any>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace argparse {

class Argument {
public:
    explicit Argument(std::string n) : name_(std::move(n)) {}

    template <typename T>
    Argument& default_value(T v) {
        value_ = std::any(std::move(v));
        return *this;
    }
    Argument& default_value(const char* v) { return default_value(std::string(v)); }

    template <typename T>
    Argument& implicit_value(T v) {
        implicit_ = std::any(std::move(v));
        is_flag_ = true;
        return *this;
    }

    template <char Shape, typename T>
    Argument& scan() {
        parse_ = [](const std::string& s) -> std::any {
            std::istringstream is(s);
            T out{};
            is >> out;
            if (is.fail()) throw std::runtime_error("bad value '" + s + "'");
            return std::any(out);
        };
        return *this;
    }

    template <typename... Ts>
    Argument& choices(Ts... cs) {
        (choices_.emplace_back(cs), ...);
        return *this;
    }
    Argument& help(const std::string&) { return *this; }

private:
    friend class ArgumentParser;
    std::string name_;
    std::any value_;
    std::any implicit_;
    bool is_flag_ = false;
    std::any (*parse_)(const std::string&) = nullptr;
    std::vector<std::string> choices_;
};

class ArgumentParser {
public:
    explicit ArgumentParser(std::string prog, std::string = "") : prog_(std::move(prog)) {}

    Argument& add_argument(const std::string& name) {
        order_.push_back(name);
        auto it = args_.emplace(name, std::make_unique<Argument>(name)).first;
        return *it->second;
    }
    Argument& add_argument(const std::string& short_name, const std::string& name) {
        Argument& a = add_argument(name);
        alias_[short_name] = name;
        return a;
    }

    void parse_args(int argc, const char* const* argv) {
        for (int i = 1; i < argc; ++i) {
            std::string key = argv[i];
            if (alias_.count(key)) key = alias_[key];
            auto it = args_.find(key);
            if (it == args_.end()) throw std::runtime_error("Unknown argument: " + key);
            Argument& a = *it->second;
            if (a.is_flag_) {
                a.value_ = a.implicit_;
                continue;
            }
            if (i + 1 >= argc) throw std::runtime_error("Missing value for " + key);
            std::string raw = argv[++i];
            if (!a.choices_.empty()) {
                bool ok = false;
                for (auto& c : a.choices_) ok = ok || (c == raw);
                if (!ok) throw std::runtime_error("Invalid choice '" + raw + "' for " + key);
            }
            a.value_ = a.parse_ ? a.parse_(raw) : std::any(raw);
        }
    }

    template <typename T>
    T get(const std::string& name) const {
        auto it = args_.find(name);
        if (it == args_.end()) throw std::runtime_error("No such argument: " + name);
        return std::any_cast<T>(it->second->value_);
    }

    friend std::ostream& operator<<(std::ostream& os, const ArgumentParser& p) {
        os << "Usage: " << p.prog_;
        for (auto& n : p.order_) os << " [" << n << "]";
        return os << "\n";
    }

private:
    std::string prog_;
    std::vector<std::string> order_;
    std::map<std::string, std::unique_ptr<Argument>> args_;
    std::map<std::string, std::string> alias_;
};

}  // namespace argparse
