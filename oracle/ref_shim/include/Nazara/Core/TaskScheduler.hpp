// oracle/ref_shim — stand-in for <Nazara/Core/TaskScheduler.hpp>: AddTask queues, WaitForTasks runs the queue on N worker
// threads (one task per chunk, the fan-out shape of Planet::GenerateChunks, Planet.cpp:394-402) and returns when all are done.
#pragma once
#include <atomic>
#include <exception>
#include <mutex>
#include <functional>
#include <thread>
#include <vector>

namespace Nz
{
	class TaskScheduler
	{
		public:
			explicit TaskScheduler(unsigned int workerCount = 0) : m_workerCount(workerCount ? workerCount : std::thread::hardware_concurrency()) {}

			template<typename F> void AddTask(F&& f) { m_tasks.emplace_back(std::forward<F>(f)); }

			void WaitForTasks()
			{
				std::atomic<std::size_t> next{ 0 };
				std::exception_ptr failure;
				std::mutex failureMutex;
				auto work = [&] {
					for (;;)
					{
						std::size_t i = next.fetch_add(1);
						if (i >= m_tasks.size())
							return;
						try
						{
							m_tasks[i]();
						}
						catch (...)
						{
							std::lock_guard<std::mutex> lock(failureMutex);
							if (!failure)
								failure = std::current_exception();
						}
					}
				};
				if (m_workerCount <= 1)
					work();
				else
				{
					std::vector<std::thread> workers;
					for (unsigned int t = 0; t < m_workerCount; ++t)
						workers.emplace_back(work);
					for (auto& w : workers)
						w.join();
				}
				m_tasks.clear();
				if (failure)
					std::rethrow_exception(failure); // e.g. Nz::SafeCast out of range inside Planet::GenerateChunk
			}

		private:
			std::vector<std::function<void()>> m_tasks;
			unsigned int m_workerCount;
	};
}
