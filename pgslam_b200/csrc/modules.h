// modules.h — host-side mirror of libpointmatcher's Registrar / Parametrizable
// (SURVEY.md §8a rows A17/A18): string-typed, validated parameters; module
// lookup by name per kind; and the YAML-configured ICP chain
// (ICPChainBase::loadFromYaml / setDefault, Localizer.hpp:70, LoopCloser.hpp:73).
#pragma once

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace pgs {

enum class Kind { DataPointsFilter, Matcher, OutlierFilter, ErrorMinimizer, TransformationChecker, Inspector, Logger, Transformation };

struct ParamDoc {
  const char* name;
  const char* doc;
  const char* def;
  const char* min;  // "" = unbounded
  const char* max;
  char type;  // 'i' integer, 'u' unsigned/bool, 'f' real, 's' string
};

using Params = std::map<std::string, std::string>;

// A created module: registered name + fully defaulted, validated parameters.
struct Module {
  Kind kind;
  std::string name;
  Params params;
  double real(const std::string& k) const;
  int64_t integer(const std::string& k) const;
  bool flag(const std::string& k) const { return integer(k) != 0; }
};

// REG(kind).create(name, params): unknown name -> PGS_INVALID_ELEMENT,
// unknown / out-of-range parameter -> PGS_INVALID_PARAMETER.
Module create_module(Kind kind, const std::string& name, const Params& params);
const std::vector<ParamDoc>* module_params(Kind kind, const std::string& name);  // nullptr if unknown
std::vector<std::string> registered_modules(Kind kind);

// ---- minimal YAML (block + flow subset that libpointmatcher configs use) ---
struct YamlNode {
  enum Type { Null, Scalar, Map, Seq } type = Null;
  std::string scalar;
  std::vector<std::pair<std::string, YamlNode>> map;  // insertion ordered
  std::vector<YamlNode> seq;
  const YamlNode* get(const std::string& key) const;
};
YamlNode parse_yaml(const std::string& text);  // throws Error(PGS_INVALID_PARAMETER) on syntax errors

// "- Name: {k: v}" / "- Name" / "Name: {..}" / "Name" -> Module
Module module_from_yaml(Kind kind, const YamlNode& node);
std::vector<Module> module_list_from_yaml(Kind kind, const YamlNode& node);

struct ChainConfig {
  std::vector<Module> reading_filters, reading_step_filters, reference_filters;
  Module matcher;
  std::vector<Module> outlier_filters;
  Module minimizer;
  std::vector<Module> checkers;
  Module inspector, logger;
};
ChainConfig chain_default();                          // ICPChainBase::setDefault
ChainConfig chain_from_yaml(const std::string& text); // ICPChainBase::loadFromYaml

}  // namespace pgs
