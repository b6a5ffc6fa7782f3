// oracle/ref_shim — stand-in for <NazaraUtils/Signal.hpp>: signal + scoped connection (slot)
#pragma once
#include <functional>
#include <memory>
#include <vector>

namespace Nz
{
	template<typename... Args> class Signal
	{
		public:
			struct SlotData
			{
				std::function<void(Args...)> callback;
				bool alive = true;
			};
			using SlotPtr = std::shared_ptr<SlotData>;

			class ConnectionGuard
			{
				public:
					ConnectionGuard() = default;
					ConnectionGuard(const ConnectionGuard&) = delete;
					ConnectionGuard(ConnectionGuard&& o) noexcept : m_slot(std::move(o.m_slot)) {}
					ConnectionGuard& operator=(ConnectionGuard&& o) noexcept { Disconnect(); m_slot = std::move(o.m_slot); return *this; }
					~ConnectionGuard() { Disconnect(); }

					template<typename F> void Connect(Signal& signal, F&& f) { Disconnect(); m_slot = signal.Connect(std::forward<F>(f)); }
					template<typename O> void Connect(Signal& signal, O* object, void (O::*method)(Args...))
					{
						Connect(signal, [object, method](Args... args) { (object->*method)(args...); });
					}
					void Disconnect()
					{
						if (m_slot)
							m_slot->alive = false;
						m_slot.reset();
					}

				private:
					SlotPtr m_slot;
			};

			Signal() = default;
			Signal(const Signal&) = delete;
			Signal(Signal&&) noexcept = default;
			Signal& operator=(Signal&&) noexcept = default;

			template<typename F> SlotPtr Connect(F&& f)
			{
				auto slot = std::make_shared<SlotData>();
				slot->callback = std::forward<F>(f);
				m_slots.push_back(slot);
				return slot;
			}

			void operator()(Args... args) const
			{
				for (std::size_t i = 0; i < m_slots.size(); ++i)
					if (m_slots[i]->alive)
						m_slots[i]->callback(args...);
			}

		private:
			std::vector<SlotPtr> m_slots;
	};
}
#define NazaraSignal(Name, ...) Nz::Signal<__VA_ARGS__> Name
#define NazaraSlot(Class, SignalName, SlotName) typename decltype(Class::SignalName)::ConnectionGuard SlotName
