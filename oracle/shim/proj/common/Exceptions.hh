// Shadows common/Exceptions.hh of the reference: the exception types the hot path throws, without Boost.Exception.
#ifndef iSAAC_COMMON_EXCEPTIONS_HH
#define iSAAC_COMMON_EXCEPTIONS_HH
#include <stdexcept>
#include <string>
namespace isaac { namespace common {
struct IsaacException : public std::runtime_error { IsaacException(const std::string &m) : std::runtime_error(m) {} IsaacException(int, const std::string &m) : std::runtime_error(m) {} };
struct InvalidParameterException : public std::logic_error { InvalidParameterException(const std::string &m) : std::logic_error(m) {} };
struct PreConditionException : public std::logic_error { PreConditionException(const std::string &m) : std::logic_error(m) {} };
struct PostConditionException : public std::logic_error { PostConditionException(const std::string &m) : std::logic_error(m) {} };
struct InvalidOptionException : public std::logic_error { InvalidOptionException(const std::string &m) : std::logic_error(m) {} };
} }
#define BOOST_THROW_EXCEPTION(e) throw (e)
#endif
