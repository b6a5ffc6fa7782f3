// oracle/ref_shim — stand-in for <NazaraUtils/FunctionRef.hpp>: non-owning reference to a callable
#pragma once
#include <cstddef>
#include <memory>
#include <type_traits>
#include <utility>

namespace Nz
{
	template<typename T> class FunctionRef;

	template<typename R, typename... Args> class FunctionRef<R(Args...)>
	{
		public:
			constexpr FunctionRef(std::nullptr_t = nullptr) noexcept : m_callback(nullptr), m_functor(nullptr) {}
			template<typename F, typename = std::enable_if_t<std::is_invocable_r_v<R, F&, Args...> && !std::is_same_v<std::decay_t<F>, FunctionRef>>>
			FunctionRef(F&& f) noexcept : m_functor(const_cast<void*>(static_cast<const void*>(std::addressof(f))))
			{
				m_callback = [](void* functor, Args... args) -> R { return (*static_cast<std::add_pointer_t<std::remove_reference_t<F>>>(functor))(std::forward<Args>(args)...); };
			}
			FunctionRef(const FunctionRef&) = default;
			FunctionRef& operator=(const FunctionRef&) = default;

			R operator()(Args... args) const { return m_callback(m_functor, std::forward<Args>(args)...); }
			explicit operator bool() const { return m_callback != nullptr; }

		private:
			R (*m_callback)(void*, Args...);
			void* m_functor;
	};
}
